/* TEST INFRASTRUCTURE: LD_PRELOAD shim that makes the reference binary believe its input files are
 * large.  SeqArc-1.6 turns on `usemodel` (train a model snapshot on the head blocks, start every
 * block from it) only when the inputs exceed ~14.07 GB (doCheckSetEncodeOpt @0x408104-0x40831f, sizes
 * from __xstat @0x40cc22); there is no command-line switch.  With SQ_FAKE_SIZE=<bytes> set, stat() of
 * any path ending in ".fq" reports that size, so the usemodel path can be pinned with small inputs. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

int __xstat(int ver, const char *path, struct stat *st) {
    static int (*real)(int, const char *, struct stat *);
    if (!real) real = (int (*)(int, const char *, struct stat *))dlsym(RTLD_NEXT, "__xstat");
    int r = real(ver, path, st);
    const char *f = getenv("SQ_FAKE_SIZE");
    size_t n = path ? strlen(path) : 0;
    if (r == 0 && f && n > 3 && strcmp(path + n - 3, ".fq") == 0) st->st_size = (off_t)atoll(f);
    return r;
}
