// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// Headless implementation of the reference's platform layer (platform.h:4-20),
// replacing platform_linux.cpp (SDL2, absent in this image; it also never
// defines MRT_ReportProgress, platform.h:10).  No window: the "title bar" the
// reference uses for its statistics (main.cpp:399-411) is captured so that the
// stock renderer can be timed by its own clock, and the close event is raised
// once the title reports completion (main.cpp:403-405 prints "Mrays/s").
#include "platform.h"
#include "main.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/resource.h>
#include <sys/syscall.h>

static char G_lastTitle[256];
static volatile bool G_traceDone = false;
bool MRT_headless_quiet = false;

const char *MRT_headless_last_title() { return G_lastTitle; }

uint64_t MRT_GetTime() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
    return (uint64_t) ts.tv_nsec + (uint64_t) ts.tv_sec * 1000000000ull;
}

float MRT_TimeDelta(uint64_t start, uint64_t stop) {
    return (float) ((stop - start) / 1000000000.0);
}

void MRT_PlatformInit() {}
void MRT_PlatformDestroy() {}

void MRT_SetWindowTitle(const char *str) {
    strncpy(G_lastTitle, str, sizeof(G_lastTitle) - 1);
    if (strstr(str, "Mrays/s")) {
        G_traceDone = true;
    }
}

void MRT_CreateWindow(uint32_t, uint32_t, uint32_t, uint32_t) {}
void MRT_DrawToWindow(const uint32_t *) {}
void MRT_ReportProgress(uint64_t, uint64_t) {}

void MRT_HandleMessages() {
    if (G_traceDone) {
        G_traceDone = false;
        MRT::WindowCallback(MRT::MRT_CLOSE);
    }
}

void MRT_DebugPrint(const char *format, ...) {
    if (MRT_headless_quiet) return;
    va_list args;
    va_start(args, format);
    vfprintf(stderr, format, args);
    va_end(args);
}

void MRT_Assert(bool cond) {
    if (!cond) {
        fprintf(stderr, "MRT_Assert failed\n");
        abort();
    }
}

void MRT_Assert(bool cond, const char *msg) {
    if (!cond) {
        fprintf(stderr, "MRT_Assert failed: %s\n", msg ? msg : "");
        abort();
    }
}

void MRT_Sleep(uint32_t ms) {
    usleep(ms * 1000u);
}

void MRT_LowerThreadPriority() {}
