/* TEST / MEASUREMENT INFRASTRUCTURE — never shipped, never linked by the product.
 *
 * est-fact-cells = the UNMODIFIED reference est-fact with DP cell counters, without touching
 * a line of its sources: the reference is compiled where it lies into a position-independent shared library
 * (oracle/_ref/libref_estfact_pic.so; its main() renamed at compile time), in which every call to a global function —
 * also from inside the defining file — goes through the PLT.  This executable defines functions of the same names: they
 * count the cells of the reference's recurrence from the arguments and forward to the real routine (dlsym RTLD_NEXT).
 * The counts are what bench.py's GCUPS are divided from when the file exists ("cells the reference computes on this
 * input"), as opposed to the cells of the jobs our own host chose to issue.
 *
 *   compute_alignment           src/compute-alignments.c:39     n*m unless the strings are identical (:48-58)
 *   K_band_edit_distance        src/compute-alignments.c:319    (2k+1)*m if 2k+1 < n else n*m; 0 if equal or |n-m| > k
 *   edit_distance               src/refine.c:51                 ls1*ls2 (inside general_refine_borders: the BORDERS cells)
 *   compute_edit_distance       src/compute-alignments.c:235    l1*l2 unless equal (through edit_distance_matrix)
 *   edit_distance_matrix        src/compute-alignments.c:210    l1*l2 (find_longest_affix, best prefix / suffix cut)
 *   compute_gap_alignment       src/refine-intron.c:560         3*n*m
 * find_longest_common_factor_dp (src/factorization-refinement.c:255) is static and cannot be counted this way.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { C_ALIGN, C_KBAND, C_EDIT, C_BORDERS, C_GAP, C_AFFIX, C_COUNT };
static const char *NAMES[C_COUNT] = {"ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX"};
static unsigned long long cells[C_COUNT], calls[C_COUNT];
static int in_borders, in_ced;
#define REAL(name) static __typeof__(&name) real; if (!real) real = (__typeof__(&name))dlsym(RTLD_NEXT, #name)

typedef struct _list *plist;
plist compute_alignment(char *est, char *gen, bool only_one) {
  REAL(compute_alignment);
  const size_t n = strlen(est), m = strlen(gen);
  if (!(n == m && strcmp(est, gen) == 0)) { cells[C_ALIGN] += (unsigned long long)n * m; ++calls[C_ALIGN]; }
  return real(est, gen, only_one);
}

bool K_band_edit_distance(char *s1, char *s2, unsigned int k, unsigned int *edit) {
  REAL(K_band_edit_distance);
  size_t n = strlen(s1), m = strlen(s2);
  if (n < m) { const size_t t = n; n = m; m = t; }
  if (!(n == m && strcmp(s1, s2) == 0) && n - m <= k) {
    cells[C_KBAND] += (2ull * k + 1 < n) ? (2ull * k + 1) * m : (unsigned long long)n * m;
    ++calls[C_KBAND];
  }
  return real(s1, s2, k, edit);
}

unsigned int *edit_distance(const char *const s1, const size_t ls1, const char *const s2, const size_t ls2) {
  REAL(edit_distance);
  const int c = in_borders ? C_BORDERS : C_EDIT;
  cells[c] += (unsigned long long)ls1 * ls2; ++calls[c];
  return real(s1, ls1, s2, ls2);
}

bool general_refine_borders(const char *const p, const size_t len_p, const size_t min_p_cut, const size_t max_p_cut, const char *const t,
                            const size_t len_t, const unsigned int max_errs, size_t *off_p, size_t *off_t1, size_t *off_t2, unsigned int *ed) {
  REAL(general_refine_borders);
  ++in_borders;
  const bool r = real(p, len_p, min_p_cut, max_p_cut, t, len_t, max_errs, off_p, off_t1, off_t2, ed);
  --in_borders;
  return r;
}

size_t compute_edit_distance(const char *const s1, const size_t l1, const char *const s2, const size_t l2) {
  REAL(compute_edit_distance);
  ++in_ced;
  const size_t r = real(s1, l1, s2, l2);
  --in_ced;
  return r;
}

size_t *edit_distance_matrix(const char *const s1, const size_t l1, const char *const s2, const size_t l2) {
  REAL(edit_distance_matrix);
  const int c = in_ced ? C_EDIT : C_AFFIX;
  cells[c] += (unsigned long long)l1 * l2; ++calls[c];
  return real(s1, l1, s2, l2);
}

plist compute_gap_alignment(char *est, char *gen, bool one, int a, int b, int c) {
  REAL(compute_gap_alignment);
  cells[C_GAP] += 3ull * strlen(est) * strlen(gen); ++calls[C_GAP];
  return real(est, gen, one, a, b, c);
}

int ref_est_fact_main(int argc, char **argv);
int main(int argc, char **argv) {
  const int rc = ref_est_fact_main(argc, argv);
  FILE *f = fopen("cells.json", "w");
  if (f) {
    fprintf(f, "{");
    for (int i = 0; i < C_COUNT; ++i) fprintf(f, "%s\"%s\": {\"cells\": %llu, \"calls\": %llu}", i ? ", " : "", NAMES[i], cells[i], calls[i]);
    fprintf(f, "}\n");
    fclose(f);
  }
  return rc;
}
