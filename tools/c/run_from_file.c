/*
 * run_from_file.c -- plain-C client of libscsb200.so: read an SCS data file, solve it on the B200,
 * print the outcome.  Counterpart of the reference's S/test/run_from_file.c (which links its own core);
 * shows the C ABI of include/scs_b200.h used without Python.
 *
 *   gcc -O2 -Iinclude tools/c/run_from_file.c -Lscs_python_b200 -lscsb200 -Wl,-rpath,$PWD/scs_python_b200 \
 *       -o tools/c/_bin/run_from_file
 *   tools/c/_bin/run_from_file tests/golden/rw_ref_mixed.bin [csv_trace_out]
 */
#include <stdio.h>
#include <stdlib.h>

#include "scs_b200.h"

int main(int argc, char **argv) {
  ScsData *d = NULL;
  ScsCone *k = NULL;
  ScsSettings *stgs = NULL;
  ScsSolution sol = {0};
  ScsInfo info;
  scs_int status;
  if (argc < 2) {
    fprintf(stderr, "usage: %s <scs data file> [csv trace]\n", argv[0]);
    return 2;
  }
  if (scs_b200_read_data(argv[1], &d, &k, &stgs) != 0) {
    fprintf(stderr, "could not read %s\n", argv[1]);
    return 3;
  }
  if (scs_b200_device_count() <= 0) {
    fprintf(stderr, "no CUDA device: this backend has no CPU fallback\n");
    scs_b200_free_data(d, k, stgs);
    return 4;
  }
  stgs->verbose = 0;
  if (argc > 2) stgs->log_csv_filename = argv[2];
  status = scs(d, k, stgs, &sol, &info);
  printf("file=%s n=%d m=%d status=%s status_val=%d iter=%d pobj=%.12e dobj=%.12e res_pri=%.3e res_dual=%.3e gap=%.3e "
         "solver=%s\n",
         argv[1], (int)d->n, (int)d->m, info.status, (int)status, (int)info.iter, info.pobj, info.dobj, info.res_pri,
         info.res_dual, info.gap, info.lin_sys_solver);
  free(sol.x);
  free(sol.y);
  free(sol.s);
  scs_b200_free_data(d, k, stgs);
  return status == SCS_SOLVED ? 0 : 1;
}
