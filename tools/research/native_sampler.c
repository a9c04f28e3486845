/* tools/research/native_sampler.c -- a SIGPROF stack sampler for finding where the REFERENCE's compiled Cython spends a locus
 * (analysis tool, not product code; cProfile cannot see into Cython modules and neither perf nor py-spy is in the image).
 *   gcc -O2 -fPIC -shared -o /tmp/native_sampler.so tools/research/native_sampler.c
 * Python side: tools/research/profile_reference_locus.py (symbolises the raw addresses with dladdr + nm). */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <dlfcn.h>

#define DEPTH 40
#define MAXS 200000
static void* g_frames[MAXS][DEPTH];
static int g_n[MAXS];
static volatile int g_count = 0;

static void on_prof(int sig) {
    (void)sig;
    int k = g_count;
    if (k >= MAXS) return;
    g_n[k] = backtrace(g_frames[k], DEPTH);
    g_count = k + 1;
}
int sampler_start(int interval_us) {
    void* warm[4]; backtrace(warm, 4);          /* loads libgcc's unwinder outside the handler */
    g_count = 0;
    struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_handler = on_prof; sa.sa_flags = SA_RESTART;
    if (sigaction(SIGPROF, &sa, NULL) != 0) return -1;
    struct itimerval it; it.it_interval.tv_sec = 0; it.it_interval.tv_usec = interval_us; it.it_value = it.it_interval;
    return setitimer(ITIMER_PROF, &it, NULL);
}
int sampler_stop(void) {
    struct itimerval it; memset(&it, 0, sizeof it);
    setitimer(ITIMER_PROF, &it, NULL);
    return g_count;
}
int sampler_depth(int k) { return k < g_count ? g_n[k] : 0; }
void* sampler_frame(int k, int d) { return g_frames[k][d]; }
/* dladdr for the Python side: object file name and load base of an address */
const char* sampler_module(void* addr, uintptr_t* base) {
    Dl_info info;
    if (!dladdr(addr, &info) || !info.dli_fname) return NULL;
    *base = (uintptr_t)info.dli_fbase;
    return info.dli_fname;
}
