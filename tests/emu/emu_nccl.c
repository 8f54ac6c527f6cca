/* emu_nccl.c -- TEST INFRASTRUCTURE: the five NCCL entry points the engine dlopens (tba_engine.cu, NcclApi), implemented over a
 * POSIX shared-memory segment so that the sharded solve (one rank per process, or one rank per host thread inside
 * tba_solve_multi) runs on a machine without GPUs under the SIMT emulator (tests/emu/cuda_emu.h).  ncclCommInitRank blocks until
 * every rank has joined, like the real one; ncclAllReduce reduces in rank order on every rank, so all ranks see the same bits (the
 * property the engine relies on for its replicated state).  A rank that waits longer than 300 s returns an error instead of hanging. */
#define _GNU_SOURCE
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclSystemError = 2, ncclInvalidArgument = 4 } ncclResult_t;
struct hdr { atomic_int arrived, generation; };
struct ncclComm { int rank, world; struct hdr* h; char* base; size_t bytes; };
typedef struct ncclComm* ncclComm_t;
#define SLOT ((size_t)16 << 20)
#define HDR 4096

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int barrier(ncclComm_t c) {
  const int gen = atomic_load(&c->h->generation);
  if (atomic_fetch_add(&c->h->arrived, 1) == c->world - 1) {
    atomic_store(&c->h->arrived, 0);
    atomic_fetch_add(&c->h->generation, 1);
    return 0;
  }
  const double t0 = now_s();
  while (atomic_load(&c->h->generation) == gen) {
    sched_yield();
    if (now_s() - t0 > 300.0) { fprintf(stderr, "emu_nccl: rank %d waited 300 s at a barrier\n", c->rank); return 1; }
  }
  return 0;
}

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  static atomic_int counter;
  memset(id, 0, sizeof *id);
  snprintf(id->internal, sizeof id->internal, "/tba_emu_nccl_%d_%lld_%d", (int)getpid(), (long long)(now_s() * 1e6), atomic_fetch_add(&counter, 1));
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* out, int world, ncclUniqueId id, int rank) {
  if (world < 1 || rank < 0 || rank >= world || id.internal[0] != '/') return ncclInvalidArgument;
  id.internal[127] = 0;
  const int fd = shm_open(id.internal, O_CREAT | O_RDWR, 0600);
  if (fd < 0) return ncclSystemError;
  const size_t bytes = HDR + (size_t)world * SLOT;
  if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); return ncclSystemError; }
  char* base = (char*)mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return ncclSystemError;
  ncclComm_t c = (ncclComm_t)calloc(1, sizeof *c);
  c->rank = rank; c->world = world; c->h = (struct hdr*)base; c->base = base; c->bytes = bytes;
  if (barrier(c)) { munmap(base, bytes); free(c); return ncclSystemError; }
  if (rank == 0) shm_unlink(id.internal);  /* every rank is mapped by now */
  *out = c;
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) {
  if (c) { munmap(c->base, c->bytes); free(c); }
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t n, int dtype, int op, ncclComm_t c, void* stream) {
  (void)stream;
  if (dtype != 8 || (op != 0 && op != 2)) return ncclInvalidArgument;  /* ncclDouble; ncclSum / ncclMax */
  const size_t cap = SLOT / sizeof(double);
  for (size_t off = 0; off < n || (n == 0 && off == 0); off += cap) {
    const size_t m = n - off < cap ? n - off : cap;
    memcpy(c->base + HDR + (size_t)c->rank * SLOT, (const double*)send + off, m * sizeof(double));
    if (barrier(c)) return ncclSystemError;
    double* dst = (double*)recv + off;
    for (size_t i = 0; i < m; ++i) {
      double acc = ((const double*)(c->base + HDR))[i];
      for (int r = 1; r < c->world; ++r) {
        const double v = ((const double*)(c->base + HDR + (size_t)r * SLOT))[i];
        acc = op == 0 ? acc + v : (v > acc ? v : acc);
      }
      dst[i] = acc;
    }
    if (barrier(c)) return ncclSystemError;
    if (n == 0) break;
  }
  return ncclSuccess;
}

const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : r == ncclInvalidArgument ? "invalid argument (emu)" : "system error (emu)"; }
