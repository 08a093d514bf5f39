// Host-side probe for the upload path: how fast can the box's cores turn fp64 rows into fp32 (read 8 B, write 4 B)?
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
int main(int argc, char** argv) {
  size_t n = (size_t)1 << 29;  // 4 GiB of doubles
  double* x = (double*)malloc(n * 8);
  float* y = (float*)malloc(n * 4);
  if (!x || !y) return 1;
  const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
  const char* e = getenv("OMP_NUM_THREADS");
  printf("online cpus %d, omp max threads %d, OMP_NUM_THREADS=%s\n", ncpu, omp_get_max_threads(), e ? e : "(unset)");
  omp_set_dynamic(0);
  omp_set_num_threads(ncpu);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) { x[i] = (double)i; y[i] = 0.f; }
  for (int t = 1; t <= ncpu; t *= 2) {
    omp_set_num_threads(t);
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) y[i] = (float)(x[i] - 1.5);
    double dt = omp_get_wtime() - t0;
    printf("threads %d: %.3f s, %.1f GB/s read, %.1f M rows(128)/s\n", t, dt, n * 8 / dt / 1e9, n / 128.0 / dt / 1e6);
  }
  printf("y[5]=%f\n", y[5]);
  return 0;
}
