/* TEST INFRASTRUCTURE — never shipped, never linked by the product.
 *
 * Thin exported wrappers around the UNMODIFIED reference DP / seeding routines so that tests and the
 * fixture generator can call the real thing through ctypes.  The reference sources are compiled where
 * they lie under $(REF); this file #includes src/factorization-refinement.c to reach its `static`
 * routines (find_longest_common_factor_dp :255, find_longest_affix :1134), the same trick the
 * reference's own unit tests use (test/refine-intron_test.c:4-12).
 */
#include "src/factorization-refinement.c"

#include "compute-alignments.h"
#include "refine.h"
#include "refine-intron.h"
#include "max-emb-graph.h"
#include "aug_suffix_tree.h"
#include "meg-simplification.h"
#include "exon-complexity.h"

/* compute-alignments.c:39 */
int ref_align(char *est, char *gen, char *out_est_aln, char *out_gen_aln, int *out_dim) {
  plist l = compute_alignment(est, gen, true);
  palignment a = (palignment)list_head(l);
  int score = a->score;
  *out_dim = a->alignment_dim;
  memcpy(out_est_aln, a->EST_alignment, a->alignment_dim + 1);
  memcpy(out_gen_aln, a->GEN_alignment, a->alignment_dim + 1);
  alignments_destroy(l);
  return score;
}

/* compute-alignments.c:319 */
int ref_kband(char *s1, char *s2, unsigned int ub, unsigned int *edit) {
  return K_band_edit_distance(s1, s2, ub, edit) ? 1 : 0;
}

/* refine.c:51 (last cell) */
unsigned int ref_edit_distance(const char *s1, size_t l1, const char *s2, size_t l2) {
  unsigned int *M = edit_distance(s1, l1, s2, l2);
  unsigned int r = M[(l1 + 1) * (l2 + 1) - 1];
  pfree(M);
  return r;
}

/* compute-alignments.c:235 */
size_t ref_compute_edit_distance(const char *s1, size_t l1, const char *s2, size_t l2) {
  return compute_edit_distance(s1, l1, s2, l2);
}

/* compute-alignments.c:246,290 */
size_t ref_best_suffix_cut(const char *s1, size_t l1, const char *s2, size_t l2, size_t *c1, size_t *c2) {
  return compute_best_suffix_cut(s1, l1, s2, l2, c1, c2);
}
size_t ref_best_prefix_cut(const char *s1, size_t l1, const char *s2, size_t l2, size_t *c1, size_t *c2) {
  return compute_best_prefix_cut(s1, l1, s2, l2, c1, c2);
}

/* refine.c:106 */
int ref_refine_borders(const char *p, size_t len_p, size_t min_p_cut, size_t max_p_cut, const char *t, size_t len_t,
                       unsigned int max_errs, size_t *off_p, size_t *off_t1, size_t *off_t2, unsigned int *ed) {
  return general_refine_borders(p, len_p, min_p_cut, max_p_cut, t, len_t, max_errs, off_p, off_t1, off_t2, ed) ? 1 : 0;
}

/* refine-intron.c:560; out_pos = factor_cut, intron_start, intron_end, intron_start_on_align, intron_end_on_align */
int ref_gap_align(char *est, char *gen, char *out_est_aln, char *out_gen_aln, int *out_pos) {
  plist l = compute_gap_alignment(est, gen, true, 0, 0, 0);
  pgap_alignment a = (pgap_alignment)list_head(l);
  int dim = a->gap_alignment_dim;
  memcpy(out_est_aln, a->EST_gap_alignment, dim + 1);
  memcpy(out_gen_aln, a->GEN_gap_alignment, dim + 1);
  out_pos[0] = a->factor_cut;
  out_pos[1] = a->intron_start;
  out_pos[2] = a->intron_end;
  out_pos[3] = a->intron_start_on_align;
  out_pos[4] = a->intron_end_on_align;
  gap_alignments_destroy(l);
  return dim;
}

/* factorization-refinement.c:255 */
void ref_lcs(const char *s1, size_t l1, const char *s2, size_t l2, size_t *occ1, size_t *occ2, size_t *len) {
  find_longest_common_factor_dp(s1, l1, s2, l2, occ1, occ2, len);
}

/* factorization-refinement.c:1134 */
int ref_longest_affix(char *est, size_t estl, char *gen, size_t genl, size_t *ecut, size_t *gcut) {
  return find_longest_affix(est, estl, gen, genl, ecut, gcut) ? 1 : 0;
}

/* refine-intron.c:376 */
int ref_burset(const char *donor, const char *acceptor) {
  char d[3] = { donor[0], donor[1], 0 }, a[3] = { acceptor[0], acceptor[1], 0 };
  return getBursetFrequency(d, a);
}

/* exon-complexity.c:50 */
double ref_dust(const char *s) { return dustScore(s); }

/* Seeding: suffix tree build + build_vertex_set exactly as main-est-fact.c:223-240 and
 * compute-est-fact.c:108-118 do, for one EST string (already strand-fixed / masked by the caller).
 * Returns the number of pairings written as (p,t,l) triples, or -(needed) if cap is too small. */
struct ref_index { LST_StringSet *set; LST_STree *tree; ppreproc_gen pg; pEST_info gen; pconfiguration cfg; };

static pconfiguration default_cfg(void) {
  pconfiguration c = PALLOC(struct _configuration);
  memset(c, 0, sizeof *c);
  c->min_factor_len = 15; c->min_intron_length = 40; c->max_intron_length = 0;
  c->min_string_depth_rate = 0.2; c->max_prefix_discarded_rate = 0.6; c->max_suffix_discarded_rate = 0.6;
  c->max_prefix_discarded = 50; c->max_suffix_discarded = 50; c->max_site_difference = 50;
  c->max_number_of_factorizations = 0; c->max_coverage_diff = 0.05; c->max_exonNUM_diff = 5;
  c->max_gapLength_diff = 20; c->retain_externals = 1; c->max_pairings_in_MEG = 80;
  c->max_freq_shortest_pairing = 0.4; c->suffpref_length_on_est = 30; c->suffpref_length_for_intron = 70;
  c->suffpref_length_on_gen = 30; c->trans_red = true; c->short_edge_comp = true;
  c->max_single_factorization_time = 900; c->complexity_threshold = 20.0;
  return c;
}

void *ref_index_create(const char *genome, int min_factor_len, double rate) {
  struct ref_index *ix = PALLOC(struct ref_index);
  ix->cfg = default_cfg();
  ix->cfg->min_factor_len = min_factor_len;
  ix->cfg->min_string_depth_rate = rate;
  ix->gen = EST_info_create();
  ix->gen->EST_seq = alloc_and_copy(genome);
  ix->set = lst_stringset_new();
  LST_String *s = PALLOC(LST_String);
  lst_string_init(s, ix->gen->EST_seq, sizeof(char), strlen(ix->gen->EST_seq));
  lst_stringset_add(ix->set, s);
  ix->tree = lst_stree_new(ix->set);
  ix->pg = PGen_create();
  preprocess_text(ix->gen, ix->pg);
  stree_preprocess(ix->tree, ix->pg, ix->cfg);
  return ix;
}

/* mode 0: vertex set only (after filters A and B); mode 1: + build_edge_set + simplify pipeline is NOT run */
long ref_seed(void *vix, const char *est_seq, int min_factor_len, int *out_ptl, long cap) {
  struct ref_index *ix = vix;
  pEST_info est = EST_info_create();
  est->EST_seq = alloc_and_copy(est_seq);
  unsigned int saved = ix->cfg->min_factor_len;
  ix->cfg->min_factor_len = min_factor_len;
  pext_array V = build_vertex_set(est, ix->tree, ix->pg, ix->cfg);
  ix->cfg->min_factor_len = saved;
  long n = 0;
  size_t sz = EA_size(V);
  for (size_t i = 1; i + 1 < sz; ++i) {
    plist l = (plist)EA_get(V, i);
    plistit it = list_first(l);
    while (listit_has_next(it)) {
      ppairing q = (ppairing)listit_next(it);
      if (n < cap) { out_ptl[3 * n] = q->p; out_ptl[3 * n + 1] = q->t; out_ptl[3 * n + 2] = q->l; }
      ++n;
    }
    listit_destroy(it);
  }
  /* V and est are leaked on purpose (test harness; destructors differ between MEG states) */
  return n <= cap ? n : -n;
}
