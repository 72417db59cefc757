/* TEST DOUBLE -- not part of the product (see fake_cuda.c).  Built as tests/hostdev/_build/libnccl.so.2 and found
 * through LD_LIBRARY_PATH by the dlopen() of libnomp's src/comm.c in the multi-rank tests of tests/test_hostdev_cpu.py.
 *
 * The five NCCL entry points libnomp uses, for ranks that are processes of one machine and whose "device" memory is
 * host memory: the unique id names a directory under NOMP_HOSTDEV_DIR, an all-reduce is one file per rank and call
 * (written under a temporary name and renamed), folded in rank order by every rank.  Small counts only. */
#define _GNU_SOURCE
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define EXPORT __attribute__((visibility("default")))

struct ncclComm {
  int rank, size;
  unsigned long long calls;
  char dir[512];
};

EXPORT const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "hostdev NCCL double: failure"; }

EXPORT ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof(*id));
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  snprintf(id->internal, sizeof(id->internal), "nccl-%ld-%ld%09ld", (long)getpid(), (long)ts.tv_sec, ts.tv_nsec);
  return ncclSuccess;
}

EXPORT ncclResult_t ncclCommInitRank(ncclComm_t *comm, int size, ncclUniqueId id, int rank) {
  const char *base = getenv("NOMP_HOSTDEV_DIR");
  if (!base || size < 1 || rank < 0 || rank >= size) return ncclInvalidArgument;
  struct ncclComm *c = calloc(1, sizeof(*c));
  c->rank = rank, c->size = size;
  snprintf(c->dir, sizeof(c->dir), "%s/%.100s", base, id.internal);
  mkdir(c->dir, 0700); /* every rank tries; EEXIST is fine */
  *comm = c;
  return ncclSuccess;
}

EXPORT ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  free(comm);
  return ncclSuccess;
}

static size_t width(ncclDataType_t dt) {
  switch (dt) {
  case ncclInt32:
  case ncclUint32:
  case ncclFloat32: return 4;
  case ncclInt64:
  case ncclUint64:
  case ncclFloat64: return 8;
  default: return 0;
  }
}

#define FOLD(T)                                                                                                        \
  for (size_t i = 0; i < count; i++) {                                                                                 \
    T a = ((T *)acc)[i], b = ((const T *)in)[i];                                                                       \
    ((T *)acc)[i] = op == ncclSum ? (T)(a + b) : op == ncclProd ? (T)(a * b) : op == ncclMin ? (b < a ? b : a) : (b > a ? b : a); \
  }

static void fold(void *acc, const void *in, size_t count, ncclDataType_t dt, ncclRedOp_t op) {
  switch (dt) {
  case ncclInt32: FOLD(int32_t) break;
  case ncclUint32: FOLD(uint32_t) break;
  case ncclInt64: FOLD(int64_t) break;
  case ncclUint64: FOLD(uint64_t) break;
  case ncclFloat32: FOLD(float) break;
  default: FOLD(double) break;
  }
}

EXPORT ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm,
                                  cudaStream_t stream) {
  (void)stream;
  const size_t bytes = count * width(dt);
  if (bytes == 0 || bytes > 4096) return ncclInvalidArgument;
  const unsigned long long call = ++comm->calls;
  char name[640], tmp[660], buf[4096], acc[4096];
  snprintf(name, sizeof(name), "%s/ar.%llu.%d", comm->dir, call, comm->rank);
  snprintf(tmp, sizeof(tmp), "%s.tmp", name);
  FILE *f = fopen(tmp, "wb");
  if (!f || fwrite(send, bytes, 1, f) != 1 || fclose(f) != 0 || rename(tmp, name) != 0) return ncclSystemError;
  for (int r = 0; r < comm->size; r++) {
    snprintf(name, sizeof(name), "%s/ar.%llu.%d", comm->dir, call, r);
    int got = 0;
    for (int tries = 0; tries < 60000 && !got; tries++) { /* up to a minute */
      f = fopen(name, "rb");
      if (f) {
        got = fread(buf, bytes, 1, f) == 1;
        fclose(f);
      }
      if (!got) {
        struct timespec ts = {0, 1000 * 1000};
        nanosleep(&ts, NULL);
      }
    }
    if (!got) return ncclSystemError;
    if (r == 0) memcpy(acc, buf, bytes);
    else fold(acc, buf, count, dt, op);
  }
  memcpy(recv, acc, bytes);
  return ncclSuccess;
}
