/* ORACLE / CPU BASELINE DRIVER (test + measurement infrastructure, never on the product path).
 *
 * Times the reference's own CPU implementation of the headline step -- tensorCRTRq then tensorCRTInvRq on one ring
 * element (lol-cpp crt.cpp:562-581) -- on every host core, without a Python interpreter in the loop:
 *
 *     ref_bench <lib.so> <symbol prefix> <tables.bin> <processes> <pairs per process> <steps> <warmup steps>
 *
 * <lib.so> is oracle/_ref/libctensor_ref.so (the UNMODIFIED reference, prefix "") or oracle/liblol_oracle.so (the C
 * restatement, prefix "lo_").  One forked process per core, not threads: the reference keeps its modulus in a global
 * (`Zq::q`, common.cpp:14), so the library is not thread-safe.  tables.bin (written by bench.py from oracle/tables.py)
 * holds exactly what the Haskell side hands to C (CPP.hs:422-442):
 *
 *     int32 npe, totm, k;  int16 pe[npe][2];  int64 qs[k];  per prime power: int64 ru[p^e * k];  the same for ruinv;
 *     int64 mhatInv[k]
 *
 * Every process owns 256 distinct uniform ring elements (the distribution of the GPU batch), checks CRTInv(CRT(x)) = x
 * once, then for each step runs <pairs> CRT + CRTInv pairs between two clock_gettime calls.  The parent prints one JSON
 * line: per step the slowest process's seconds (the step's wall time on a machine with that many cores).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

typedef struct { int16_t prime, exponent; } PrimeExponent;
typedef void (*crt_fn)(int16_t, int64_t*, int32_t, PrimeExponent*, int16_t, int64_t**, int64_t*);
typedef void (*crtinv_fn)(int16_t, int64_t*, int32_t, PrimeExponent*, int16_t, int64_t**, int64_t*, int64_t*);

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

#define NELEM 256

int main(int argc, char** argv)
{
  if (argc != 8) { fprintf(stderr, "usage: ref_bench lib.so prefix tables.bin procs pairs steps warmup\n"); return 2; }
  const char* libpath = argv[1];
  const char* prefix = argv[2];
  const int procs = atoi(argv[4]), pairs = atoi(argv[5]), steps = atoi(argv[6]), warm = atoi(argv[7]);
  if (procs < 1 || procs > 1024 || pairs < 1 || steps < 1 || steps > 4096 || warm < 0) { fprintf(stderr, "ref_bench: bad counts\n"); return 2; }

  FILE* f = fopen(argv[3], "rb");
  if (!f) { perror("ref_bench: tables"); return 2; }
  int32_t hdr[3];
  if (fread(hdr, sizeof(int32_t), 3, f) != 3) { fprintf(stderr, "ref_bench: short tables file\n"); return 2; }
  const int npe = hdr[0], totm = hdr[1], k = hdr[2];
  if (npe < 1 || npe > 16 || totm < 1 || k < 1 || k > 16) { fprintf(stderr, "ref_bench: bad header\n"); return 2; }
  PrimeExponent pe[16];
  int64_t qs[16], mh[16];
  int64_t *ru[16], *rui[16];
  int ok = fread(pe, sizeof(PrimeExponent), (size_t)npe, f) == (size_t)npe && fread(qs, sizeof(int64_t), (size_t)k, f) == (size_t)k;
  for (int dir = 0; dir < 2 && ok; dir++)
    for (int i = 0; i < npe && ok; i++) {
      const size_t cnt = (size_t)ipow(pe[i].prime, pe[i].exponent) * (size_t)k;
      int64_t* t = (int64_t*)malloc(cnt * sizeof(int64_t));
      ok = t && fread(t, sizeof(int64_t), cnt, f) == cnt;
      (dir ? rui : ru)[i] = t;
    }
  ok = ok && fread(mh, sizeof(int64_t), (size_t)k, f) == (size_t)k;
  fclose(f);
  if (!ok) { fprintf(stderr, "ref_bench: short tables file\n"); return 2; }

  void* lib = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "ref_bench: %s\n", dlerror()); return 2; }
  char name[128];
  snprintf(name, sizeof name, "%stensorCRTRq", prefix);
  crt_fn crt = (crt_fn)dlsym(lib, name);
  snprintf(name, sizeof name, "%stensorCRTInvRq", prefix);
  crtinv_fn crtinv = (crtinv_fn)dlsym(lib, name);
  if (!crt || !crtinv) { fprintf(stderr, "ref_bench: symbols not found in %s\n", libpath); return 2; }

  const int total = warm + steps;
  int (*pipes)[2] = malloc(sizeof(int[2]) * (size_t)procs);
  for (int p = 0; p < procs; p++) {
    if (pipe(pipes[p]) != 0) { perror("pipe"); return 2; }
    pid_t pid = fork();
    if (pid < 0) { perror("fork"); return 2; }
    if (pid == 0) {
      close(pipes[p][0]);
      const size_t esz = (size_t)totm * (size_t)k;
      int64_t* elems = (int64_t*)malloc(NELEM * esz * sizeof(int64_t));
      int64_t* y = (int64_t*)malloc(esz * sizeof(int64_t));
      uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(p + 1);
      for (size_t i = 0; i < NELEM * esz; i++) {      // xorshift64*, uniform in [0, q_limb)
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        elems[i] = (int64_t)(((s * 0x2545F4914F6CDD1Dull) >> 11) % (uint64_t)qs[i % (size_t)k]);
      }
      memcpy(y, elems, esz * sizeof(int64_t));
      crt((int16_t)k, y, totm, pe, (int16_t)npe, ru, qs);
      crtinv((int16_t)k, y, totm, pe, (int16_t)npe, rui, mh, qs);
      if (memcmp(y, elems, esz * sizeof(int64_t)) != 0) { fprintf(stderr, "ref_bench: CRTInv(CRT(x)) != x\n"); _exit(3); }
      double* secs = (double*)malloc(sizeof(double) * (size_t)total);
      for (int st = 0; st < total; st++) {
        const double t0 = now_s();
        for (int i = 0; i < pairs; i++) {
          memcpy(y, elems + (size_t)(i & (NELEM - 1)) * esz, esz * sizeof(int64_t));      // the Haskell side thaws a copy too (CPP.hs:333-337)
          crt((int16_t)k, y, totm, pe, (int16_t)npe, ru, qs);
          crtinv((int16_t)k, y, totm, pe, (int16_t)npe, rui, mh, qs);
        }
        secs[st] = now_s() - t0;
      }
      ssize_t w = write(pipes[p][1], secs, sizeof(double) * (size_t)total);
      _exit(w == (ssize_t)(sizeof(double) * (size_t)total) ? 0 : 4);
    }
    close(pipes[p][1]);
  }
  double* worst = (double*)calloc((size_t)total, sizeof(double));
  double* sum = (double*)calloc((size_t)total, sizeof(double));
  double* buf = (double*)malloc(sizeof(double) * (size_t)total);
  int failed = 0;
  for (int p = 0; p < procs; p++) {
    size_t got = 0;
    while (got < sizeof(double) * (size_t)total) {
      ssize_t r = read(pipes[p][0], (char*)buf + got, sizeof(double) * (size_t)total - got);
      if (r <= 0) break;
      got += (size_t)r;
    }
    if (got != sizeof(double) * (size_t)total) { failed = 1; continue; }
    for (int st = 0; st < total; st++) { if (buf[st] > worst[st]) worst[st] = buf[st]; sum[st] += buf[st]; }
  }
  for (int p = 0; p < procs; p++) { int status = 0; wait(&status); if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) failed = 1; }
  if (failed) { fprintf(stderr, "ref_bench: a worker failed\n"); return 1; }
  printf("{\"procs\": %d, \"pairs\": %d, \"steps\": %d, \"warmup\": %d, \"step_seconds\": [", procs, pairs, steps, warm);
  for (int st = warm; st < total; st++) printf("%s%.6f", st > warm ? ", " : "", worst[st]);
  printf("], \"mean_process_seconds\": [");
  for (int st = warm; st < total; st++) printf("%s%.6f", st > warm ? ", " : "", sum[st] / procs);
  printf("]}\n");
  return 0;
}
