/* TEST INFRASTRUCTURE — the MEG core the device runs (pintron_b200/csrc/meg_core.h), compiled for the host as a checker:
 * tests/test_gpu_parity.py compares the record the GPU returns for a PC_SEED_BUILD_MEG job with this one on the same
 * triples (the triples themselves are checked against the oracle port).  The core as a whole is pinned against the
 * reference through the byte-parity of megs.txt / meg-edges.txt on the regression cases (tests/test_host_parity.py). */
#include <stdlib.h>
#include "pintron_cuda.h"
#include "../../pintron_b200/csrc/meg_core.h"

/* returns the number of int32 words written to out; -needed when cap_words is too small; INT64_MIN on a core error */
long long meg_host_record(const int *tri, int ntri, int est_len, int l, const pc_meg_cfg *cfg, int32_t *out, long long cap_words) {
  for (long long nints = 1 << 14;; nints *= 2) {
    int *mem = malloc(sizeof(int) * (size_t)nints);
    mg_graph g;
    mg_init(&g, mem, nints, tri, ntri);
    const int retry = g.err ? 0 : mg_build(&g, est_len, l, cfg);
    if (g.err == MG_E_SCRATCH) { free(mem); continue; }
    long long words = INT64_MIN;
    if (!g.err) {
      words = mg_record_words(&g);
      if (words > cap_words) words = -words; else mg_write_record(&g, retry, out);
    }
    free(mem);
    return words;
  }
}
