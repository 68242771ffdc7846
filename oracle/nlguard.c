/* TEST INFRASTRUCTURE: LD_PRELOAD shim for running the reference binary with -l (lossy).
 * EncapFqzComp::rblock @0x426c10 scans for a '\n' that the caller never writes after the block's
 * quality bytes and overruns the heap; padding big allocations with '\n' gives it the end-of-data
 * terminator it expects. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <string.h>
static void *(*real_malloc)(size_t);
void *malloc(size_t n) {
    if (!real_malloc) real_malloc = (void *(*)(size_t))dlsym(RTLD_NEXT, "malloc");
    if (n >= (1u << 20)) {
        char *p = (char *)real_malloc(n + 64);
        if (p) memset(p + n, '\n', 64);
        return p;
    }
    return real_malloc(n);
}
