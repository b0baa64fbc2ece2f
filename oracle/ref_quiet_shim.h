// Force-included before the reference's streamer.cpp only (see Makefile): routes the
// putchar() inside printf_nano (streamer.cpp:36) to the harness so the oracle can be silenced.
// Standard headers are pulled in first so <cstdio>'s own "#undef putchar" has already run.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <mutex>
extern "C" int efref_putchar(int c);
#define putchar efref_putchar
