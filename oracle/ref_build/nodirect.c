/* nodirect.c -- TEST INFRASTRUCTURE ONLY (our code, not the reference's).
 * LD_PRELOAD shim for tools/ref_engine_harness.py: the reference's tensor store opens its files with O_DIRECT
 * (core/aio/archer_aio_utils.cpp:16-25) and aborts the process when the filesystem refuses it (tmpfs, some overlay
 * mounts).  When the GPU box offers no O_DIRECT-capable directory the harness re-executes itself with this library
 * preloaded: open()/open64()/openat() are forwarded to libc with the O_DIRECT bit cleared.  Nothing else changes. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <stdarg.h>
#include <sys/types.h>

typedef int (*open_fn)(const char*, int, ...);
typedef int (*openat_fn)(int, const char*, int, ...);

static int forward_open(const char* sym, const char* path, int flags, mode_t mode) {
  open_fn real = (open_fn)dlsym(RTLD_NEXT, sym);
  return real(path, flags & ~O_DIRECT, mode);
}

int open(const char* path, int flags, ...) {
  mode_t mode = 0;
  if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); mode = (mode_t)va_arg(ap, int); va_end(ap); }
  return forward_open("open", path, flags, mode);
}
int open64(const char* path, int flags, ...) {
  mode_t mode = 0;
  if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); mode = (mode_t)va_arg(ap, int); va_end(ap); }
  return forward_open("open64", path, flags, mode);
}
int openat(int dirfd, const char* path, int flags, ...) {
  mode_t mode = 0;
  if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); mode = (mode_t)va_arg(ap, int); va_end(ap); }
  openat_fn real = (openat_fn)dlsym(RTLD_NEXT, "openat");
  return real(dirfd, path, flags & ~O_DIRECT, mode);
}
