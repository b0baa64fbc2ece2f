// Force-included before the reference's video.cpp (see Makefile): the simulator branch of
// video.cpp uses memcpy and vTaskDelay without declaring them (SURVEY.md §8c build recipe).
#include <cstring>
static inline void vTaskDelay(int) {}
