/* refine_intron.c — splice-site refinement of one intron (reference src/refine-intron.c).
 *
 * refine_intron :47-265 aligns the EST window around the cut against donor-suffix | intron-prefix | intron-suffix |
 * acceptor-prefix with the three-plane gap alignment (a PC_OP_GAP job, src/refine-intron.c:560-890), then moves the
 * cut to the nearest GT-AG, else GC-AG, else best Burset dinucleotide pair.  The four shift heuristics
 * (Shift_right_to_left_1/2 :992,:1214, Shift_left_to_right_1/2 :1429,:1645) differ only in direction and in how a
 * candidate is accepted, so they are one routine here; every small edit distance they need is computed up front as
 * PC_OP_EDIT jobs of ONE batch (the reference evaluates them lazily, the values are the same).
 * Arithmetic quirks that decide results are kept: the `unsigned` error of variant 1 that may wrap, the `int` error
 * of variant 2 that may go negative, l_substr growing when the extension window starts before column 0.
 */
#include "ef.h"

/* strcmp(pt, pat) == 0 for the two-letter windows the splice-site scans compare at every alignment column (a libc call per
 * column was 7 % of the host profile): pt = {c0, c1, NUL}, pat = a NUL-terminated dinucleotide */
static inline bool eq2(const char *pt, const char *pat) {
  if (pt[0] != pat[0]) return false;
  if (pt[0] == 0) return true;
  if (pt[1] != pat[1]) return false;
  return pt[1] == 0 || pat[2] == 0;
}

/* ---- Burset frequencies (getBursetFrequency :376-556, pinned by reference test/refine-intron_test.c:148-922) -- */
static const struct { char d[3], a[3]; int f; } BURSET[] = {
  {"AA","AG",1},{"AA","AT",1},{"AA","GT",1},{"AC","CC",1},{"AG","AC",1},{"AG","AG",5},{"AG","CT",2},{"AG","GC",1},
  {"AG","TG",2},{"AT","AA",1},{"AT","AC",8},{"AT","AG",7},{"AT","AT",2},{"AT","GC",1},{"AT","GT",1},{"CA","AG",1},
  {"CA","TT",1},{"CC","AG",2},{"CG","AG",1},{"CG","CA",1},{"CT","AC",2},{"CT","CA",1},{"GA","AG",8},{"GA","GT",1},
  {"GA","TC",1},{"GA","TG",1},{"GC","AG",126},{"GC","GG",1},{"GC","TA",1},{"GG","AC",1},{"GG","AG",11},{"GG","CA",1},
  {"GG","GA",2},{"GG","TC",2},{"GT","AG",200},{"GT","AC",4},{"GT","AT",2},{"GT","CA",9},{"GT","CG",4},{"GT","CT",3},
  {"GT","GC",1},{"GT","GG",10},{"GT","GT",1},{"GT","TA",7},{"GT","TC",2},{"GT","TG",8},{"GT","TT",2},{"TA","AG",6},
  {"TA","CG",1},{"TA","TC",1},{"TC","AG",1},{"TC","GG",1},{"TG","AC",1},{"TG","AG",7},{"TG","GG",2},{"TT","AG",5},
  {"TT","AT",1},{"TT","GG",1}};

static char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

/* donor / acceptor: NUL-terminated strings; anything that is not exactly two letters scores 0 */
int burset_freq(const char *donor, const char *acceptor) {
  if (!donor[0] || !donor[1] || donor[2] || !acceptor[0] || !acceptor[1] || acceptor[2]) return 0;
  const char d0 = up(donor[0]), d1 = up(donor[1]), a0 = up(acceptor[0]), a1 = up(acceptor[1]);
  for (size_t i = 0; i < sizeof BURSET / sizeof BURSET[0]; ++i)
    if (BURSET[i].d[0] == d0 && BURSET[i].d[1] == d1 && BURSET[i].a[0] == a0 && BURSET[i].a[1] == a1) return BURSET[i].f;
  return 0;
}

int burset_adaptor(const char *t, size_t cut1, size_t cut2) {      /* getBursetFrequency_adaptor :362-374 */
  if (cut2 < 2) return 0;
  char d[3] = {t[cut1], 0, 0}, a[3] = {t[cut2 - 2], t[cut2 - 1], 0};
  if (d[0]) d[1] = t[cut1 + 1];
  return burset_freq(d, a);
}

/* real_substring (src/util.c:137-156) as a view: a negative index shortens the piece, the end of s truncates it */
typedef struct sv { const char *p; int len; } sv;
static sv substr(const char *s, int slen, int index, int length) {
  sv v = {s, 0};
  if (index < 0) { length += index; index = 0; }
  if (length <= 0 || index >= slen) { v.p = s + (index < slen ? index : slen); return v; }
  v.p = s + index;
  v.len = MIN2(length, slen - index);
  return v;
}

static int check_burset(const char *g, int glen, int donor_left, int acceptor_right) {   /* Check_Burset_patterns :346-360 */
  sv d = substr(g, glen, donor_left + 1, 2), a = substr(g, glen, acceptor_right - 2, 2);
  char ds[3] = {0, 0, 0}, as[3] = {0, 0, 0};
  memcpy(ds, d.p, (size_t)d.len); memcpy(as, a.p, (size_t)a.len);
  return burset_freq(ds, as);
}

/* ---- the alignment as the heuristics see it ----------------------------------------------------------------- */
typedef struct gapaln {
  char *est, *gen;            /* rows, NUL-terminated, 16 readable zero bytes before [0] */
  int dim, factor_cut, intron_start, intron_end, is_on_align, ie_on_align;
  int new_acceptor_factor_left, new_donor_right_on_gen, new_acceptor_left_on_gen;
} gapaln;

static void find_AG_after_on_the_right(const gapaln *A, int init, int *cut_on_align, int *gen_cut, int *est_cut) {   /* :892-940 */
  *cut_on_align = *gen_cut = *est_cut = -1;
  size_t index = (size_t)(init - 2);
  bool stop = false;
  while (!stop && index < (size_t)A->dim - 1) {
    while (A->gen[index] == '-') ++index;
    char pt[3];
    pt[0] = A->gen[index];
    ++index;
    while (A->gen[index] == '-') ++index;
    pt[1] = A->gen[index];
    pt[2] = 0;
    stop = eq2(pt, "AG");
  }
  if (!stop) return;
  int cg = 0, ce = 0;
  *cut_on_align = (int)index + 1;
  for (size_t i = (size_t)(A->ie_on_align + 1); i <= index; ++i) {
    if (A->gen[i] != '-') ++cg;
    if (A->est[i] != '-') ++ce;
  }
  *gen_cut = cg; *est_cut = ce;
}

static void find_ACCEPTOR_before_on_the_left(const gapaln *A, int init, int *cut_on_align, int *gen_cut, int *est_cut,
                                             const char *acc) {                                    /* :942-990 */
  *cut_on_align = *gen_cut = *est_cut = -1;
  int index = init + 2;
  bool stop = false;
  while (!stop && index > 0) {
    while (A->gen[index] == '-') --index;
    char pt[3];
    pt[1] = A->gen[index];
    --index;
    while (index >= 0 && A->gen[index] == '-') --index;
    pt[0] = index < 0 ? 0 : A->gen[index];
    pt[2] = 0;
    if (eq2(pt, acc)) stop = true;
  }
  if (!stop) return;
  int cg = 0, ce = 0;
  *cut_on_align = index - 1;
  for (int i = A->is_on_align - 1; i >= index; --i) {
    if (A->gen[i] != '-') ++cg;
    if (A->est[i] != '-') ++ce;
  }
  *gen_cut = cg; *est_cut = ce;
}

static void find_ACCEPTOR_after_on_the_left(const gapaln *A, int init, int *gen_sub, const char *acc) {   /* :1867-1890 */
  *gen_sub = -1;
  int index = init;
  bool stop = false;
  while (!stop && index < A->ie_on_align) {
    char pt[3];
    pt[0] = A->gen[index];
    ++index;
    pt[1] = A->gen[index];
    pt[2] = 0;
    if (eq2(pt, acc)) stop = true;
  }
  if (stop) *gen_sub = index - A->is_on_align - 1;
}

static void find_AG_before_on_the_right(const gapaln *A, int init, int *gen_sub) {                   /* :1958-1980 */
  *gen_sub = -1;
  int index = init;
  bool stop = false;
  while (!stop && index > A->is_on_align) {
    char pt[3];
    pt[1] = A->gen[index];
    --index;
    pt[0] = A->gen[index];
    pt[2] = 0;
    if (eq2(pt, "AG")) stop = true;
  }
  if (stop) *gen_sub = A->ie_on_align - index - 1;
}

/* Get_est_substring_from_alignment / Get_genomic_substring_from_alignment (:1892-1956): the non-gap letters of one
 * row inside [init, init+length), and the number of differing columns there */
static char *row_letters(ef_task *T, const gapaln *A, bool est_row, int init, int length, int *error) {
  if (init < 0 || init >= A->dim) return NULL;
  const int actual = A->dim - init < length ? A->dim - init : length;
  char *out = ar_alloc(&T->ar, (size_t)actual + 1);
  const char *row = est_row ? A->est : A->gen;
  int n = 0, herr = 0;
  for (int x = init; x < init + actual; ++x) {
    if (row[x] != '-') out[n++] = row[x];
    if (A->gen[x] != A->est[x]) ++herr;
  }
  *error = herr;
  return out;
}

static char *cat2(ef_task *T, const char *a, int la, const char *b, int lb) {
  char *s = ar_alloc(&T->ar, (size_t)la + (size_t)lb + 1);
  memcpy(s, a, (size_t)la); memcpy(s + la, b, (size_t)lb);
  return s;
}

#define CYCLES 2
typedef struct shiftplan {
  bool r2l, variant2;
  int gen_cut[CYCLES], est_cut[CYCLES], gen_sub[CYCLES];
  sv cut_factor[CYCLES], match_str[CYCLES], prev_match[CYCLES];
  bool has_cut[CYCLES], has_match[CYCLES];
  char *ext_cut[CYCLES], *ext_match[CYCLES];
  int ext_error;
  int h_prev[CYCLES], h_pair[CYCLES][CYCLES];
} shiftplan;

/* the candidate cuts of one shift routine and the PC_OP_EDIT jobs its decision needs */
static void shift_prepare(ef_task *T, shiftplan *S, const gapaln *A, const char *est, int elen, const char *gen, int glen,
                          bool r2l, bool variant2, const char *donor_pt) {
  memset(S, 0, sizeof *S);
  S->r2l = r2l; S->variant2 = variant2; S->ext_error = -1;
  char *ext_e, *ext_g;
  int init_right, init_left, cut_on_align = 0;
  if (r2l) {
    init_right = A->ie_on_align + 1; init_left = A->is_on_align;
    int l_sub = 8, start = A->is_on_align - l_sub;
    if (start < 0) { l_sub = l_sub - start; start = 0; }
    ext_e = row_letters(T, A, true, start, l_sub, &S->ext_error);
    ext_g = row_letters(T, A, false, start, l_sub, &S->ext_error);
  } else {
    init_right = A->ie_on_align; init_left = A->is_on_align - 1;
    ext_e = row_letters(T, A, true, A->ie_on_align + 1, 8, &S->ext_error);
    ext_g = row_letters(T, A, false, A->ie_on_align + 1, 8, &S->ext_error);
  }
  for (int i = 0; i < CYCLES; ++i) {
    if (r2l) find_AG_after_on_the_right(A, init_right, &cut_on_align, &S->gen_cut[i], &S->est_cut[i]);
    else find_ACCEPTOR_before_on_the_left(A, init_left, &cut_on_align, &S->gen_cut[i], &S->est_cut[i], donor_pt);
    if (S->est_cut[i] > -1) {
      S->has_cut[i] = true;
      if (r2l) {
        S->prev_match[i] = substr(gen, glen, A->new_acceptor_left_on_gen, S->gen_cut[i]);
        S->cut_factor[i] = substr(est, elen, A->new_acceptor_factor_left, S->est_cut[i]);
        init_right = cut_on_align + 1;
      } else {
        S->prev_match[i] = substr(gen, glen, A->new_donor_right_on_gen - S->gen_cut[i] + 1, S->gen_cut[i]);
        S->cut_factor[i] = substr(est, elen, A->new_acceptor_factor_left - S->est_cut[i], S->est_cut[i]);
        init_left = cut_on_align - 1;
      }
      if (S->ext_error > 0 && ext_e)
        S->ext_cut[i] = r2l ? cat2(T, ext_e, (int)strlen(ext_e), S->cut_factor[i].p, S->cut_factor[i].len)
                            : cat2(T, S->cut_factor[i].p, S->cut_factor[i].len, ext_e, (int)strlen(ext_e));
    }
    if (r2l) find_ACCEPTOR_after_on_the_left(A, init_left, &S->gen_sub[i], donor_pt);
    else find_AG_before_on_the_right(A, init_right, &S->gen_sub[i]);
    if (S->gen_sub[i] > -1) {
      S->has_match[i] = true;
      if (r2l) {
        S->match_str[i] = substr(gen, glen, A->new_donor_right_on_gen + 1, S->gen_sub[i]);
        init_left = A->is_on_align + S->gen_sub[i] + 1;
      } else {
        S->match_str[i] = substr(gen, glen, A->new_acceptor_left_on_gen - S->gen_sub[i], S->gen_sub[i]);
        init_right = A->ie_on_align - S->gen_sub[i] - 1;
      }
      if (S->has_cut[i] && S->ext_error > 0 && ext_g)
        S->ext_match[i] = r2l ? cat2(T, ext_g, (int)strlen(ext_g), S->match_str[i].p, S->match_str[i].len)
                              : cat2(T, S->match_str[i].p, S->match_str[i].len, ext_g, (int)strlen(ext_g));
    }
  }
  for (int i = 0; i < CYCLES; ++i) {
    S->h_prev[i] = -1;
    if (!variant2 && S->has_cut[i])
      S->h_prev[i] = dp_push(PC_OP_EDIT, S_(S->cut_factor[i].p, S->cut_factor[i].len), S_(S->prev_match[i].p, S->prev_match[i].len), 0, 0, 0, 0);
    for (int j = 0; j < CYCLES; ++j) {
      S->h_pair[i][j] = -1;
      if (S->ext_cut[i] && S->ext_match[j])
        S->h_pair[i][j] = dp_push(PC_OP_EDIT, S_(S->ext_cut[i], (int)strlen(S->ext_cut[i])), S_(S->ext_match[j], (int)strlen(S->ext_match[j])), 0, 0, 0, 0);
      else if (S->has_cut[i] && S->has_match[j])
        S->h_pair[i][j] = dp_push(PC_OP_EDIT, S_(S->cut_factor[i].p, S->cut_factor[i].len), S_(S->match_str[j].p, S->match_str[j].len), 0, 0, 0, 0);
    }
  }
}

/* the acceptance loops (:1137-1180 for variant 1, :1360-1405 for variant 2), fed with the batch results */
static bool shift_decide(const shiftplan *S, const gapaln *A, int *donor_right, int *acceptor_left, int *factor_left) {
  const int sg = S->r2l ? 1 : -1;
  bool stop = false;
  if (!S->variant2) {
    unsigned error = 1000, edit_prev;
    for (int i = 0; i < CYCLES && !stop; ++i)
      for (int j = 0; j < CYCLES && !stop; ++j) {
        if (S->has_cut[i] && S->has_match[j]) {
          edit_prev = (unsigned)dp_res(S->h_prev[i])[1];
          if (edit_prev <= 5) {
            const unsigned ed = (unsigned)dp_res(S->h_pair[i][j])[1];
            error = (S->ext_cut[i] && S->ext_match[j]) ? ed - edit_prev - (unsigned)S->ext_error : ed - edit_prev;
          }
        }
        if (error <= 1) {
          *factor_left = A->new_acceptor_factor_left + sg * S->est_cut[i];
          if (S->r2l) { *donor_right = A->new_donor_right_on_gen + S->gen_sub[j]; *acceptor_left = A->new_acceptor_left_on_gen + S->gen_cut[i]; }
          else { *donor_right = A->new_donor_right_on_gen - S->gen_cut[i]; *acceptor_left = A->new_acceptor_left_on_gen - S->gen_sub[j]; }
          stop = true;
        }
      }
  } else {
    int error = 1000, edit;
    for (int i = 0; i < CYCLES && !stop; ++i)
      for (int j = 0; j < CYCLES && !stop; ++j) {
        if (S->ext_cut[i] && S->ext_match[j]) edit = (int)((unsigned)dp_res(S->h_pair[i][j])[1] - (unsigned)S->ext_error);
        else if (S->has_cut[i] && S->has_match[j]) edit = dp_res(S->h_pair[i][j])[1];
        else edit = 1000;
        if (edit < error) {
          error = edit;
          *factor_left = A->new_acceptor_factor_left + sg * S->est_cut[i];
          if (S->r2l) { *donor_right = A->new_donor_right_on_gen + S->gen_sub[j]; *acceptor_left = A->new_acceptor_left_on_gen + S->gen_cut[i]; }
          else { *donor_right = A->new_donor_right_on_gen - S->gen_cut[i]; *acceptor_left = A->new_acceptor_left_on_gen - S->gen_sub[j]; }
        }
        if (error == 0) stop = true;
      }
  }
  return stop;
}

/* Try_Burset_after_match (:267-343): slide the cut while EST and genome keep matching, keep the best Burset pair */
static void try_burset(const char *est, int elen, const char *gen, int glen, int *factor_left, int *donor_right, int *acceptor_left,
                       int donor_factor_left, int acceptor_factor_right) {
  int fl = *factor_left, al = *acceptor_left, dr = *donor_right;
  int u_fl = fl, u_al = al, u_dr = dr, freq = 0;
  bool right_to_left = false, stop = false;
  while (!stop && est[fl] == gen[al] && fl > donor_factor_left + 1) {
    if (fl == 0 || dr == -1) stop = true;
    else {
      const int f = check_burset(gen, glen, dr, al);
      if (f > freq) { freq = f; u_fl = fl; u_al = al; u_dr = dr; }
      --fl; --dr; --al;
    }
  }
  fl = *factor_left; al = *acceptor_left + 1; dr = *donor_right + 1;
  stop = false;
  while (!stop && est[fl] == gen[dr] && fl < acceptor_factor_right) {
    if (fl == elen || al == glen) stop = true;
    else {
      const int f = check_burset(gen, glen, dr, al);
      if (f > freq) { freq = f; u_fl = fl; u_al = al; u_dr = dr; right_to_left = true; }
      ++fl; ++dr; ++al;
    }
  }
  if (right_to_left) u_fl += 1;
  *factor_left = u_fl; *donor_right = u_dr; *acceptor_left = u_al;
}

bool refine_intron(ef_task *T, const ef_seq *est_info, ef_factor *donor, ef_factor *acceptor, bool first_intron) {
  const ef_config *c = T->cfg;
  const char *est = est_info->seq, *gen = T->gen->seq;
  const int elen = est_info->len, glen = T->gen->len;
  const int on_est = c->suffpref_length_on_est, for_intron = c->suffpref_length_for_intron, on_gen = c->suffpref_length_on_gen;

  int ds_left_gen = donor->gs;
  if (donor->ge - on_gen + 1 >= ds_left_gen) ds_left_gen = donor->ge - on_gen + 1;
  const sv ds_gen = substr(gen, glen, ds_left_gen, donor->ge - ds_left_gen + 1);
  int ds_left_est = donor->es;
  if (donor->ee - on_est + 1 >= ds_left_est) ds_left_est = donor->ee - on_est + 1;
  const sv ds_est = substr(est, elen, ds_left_est, donor->ee - ds_left_est + 1);
  int ap_right_gen = acceptor->ge;
  if (acceptor->gs + on_gen - 1 <= ap_right_gen) ap_right_gen = acceptor->gs + on_gen - 1;
  const sv ap_gen = substr(gen, glen, acceptor->gs, ap_right_gen - acceptor->gs + 1);
  int ap_right_est = acceptor->ee;
  if (acceptor->es + on_est - 1 <= ap_right_est) ap_right_est = acceptor->es + on_est - 1;
  const sv ap_est = substr(est, elen, acceptor->es, ap_right_est - acceptor->es + 1);
  sv gap_est = {NULL, 0};
  if (donor->ee != acceptor->es - 1) gap_est = substr(est, elen, donor->ee + 1, acceptor->es - donor->ee - 1);
  const sv ipre = substr(gen, glen, donor->ge + 1, for_intron), isuf = substr(gen, glen, acceptor->gs - for_intron, for_intron);

  const int n = ds_est.len + gap_est.len + ap_est.len, m = ds_gen.len + ipre.len + isuf.len + ap_gen.len;
  char *seq_est = ar_alloc(&T->ar, (size_t)n + 1), *seq_gen = ar_alloc(&T->ar, (size_t)m + 1);
  memcpy(seq_est, ds_est.p, (size_t)ds_est.len);
  if (gap_est.len) memcpy(seq_est + ds_est.len, gap_est.p, (size_t)gap_est.len);
  memcpy(seq_est + ds_est.len + gap_est.len, ap_est.p, (size_t)ap_est.len);
  memcpy(seq_gen, ds_gen.p, (size_t)ds_gen.len);
  memcpy(seq_gen + ds_gen.len, ipre.p, (size_t)ipre.len);
  memcpy(seq_gen + ds_gen.len + ipre.len, isuf.p, (size_t)isuf.len);
  memcpy(seq_gen + ds_gen.len + ipre.len + isuf.len, ap_gen.p, (size_t)ap_gen.len);
  const int deleted_intron_dim = acceptor->gs - donor->ge - 1 - 2 * for_intron;

  const int h = dp_push(PC_OP_GAP, S_(seq_est, n), S_(seq_gen, m), 0, 0, 0, 0);
  dp_wait();
  const int32_t *r = dp_res(h);
  ef_aln rows = aln_from_ops(T, dp_var(h), r[1], seq_est, seq_gen);
  gapaln A;
  A.est = rows.est; A.gen = rows.gen; A.dim = r[1];
  A.factor_cut = r[2]; A.intron_start = r[3]; A.intron_end = r[4]; A.is_on_align = r[5]; A.ie_on_align = r[6];
  A.new_acceptor_factor_left = ds_left_est + A.factor_cut;
  A.new_donor_right_on_gen = ds_left_gen + A.intron_start - 1;
  A.new_acceptor_left_on_gen = ds_left_gen + A.intron_end + deleted_intron_dim + 1;

  if (A.new_acceptor_factor_left == donor->es) {
    if (first_intron) { acceptor->es = A.new_acceptor_factor_left; acceptor->gs = A.new_acceptor_left_on_gen; return true; }
    return false;
  }
  if (A.new_acceptor_left_on_gen - A.new_donor_right_on_gen < c->min_intron_length) return false;
  if (abs(A.new_donor_right_on_gen - donor->ge) > 20 || abs(A.new_acceptor_left_on_gen - acceptor->gs) > 20) return false;

  int lcut = 0, lgen = 0, lest = 0, rcut = 0, rgen = 0, rest = 0;
  find_ACCEPTOR_before_on_the_left(&A, A.is_on_align - 1, &lcut, &lgen, &lest, "GT");
  find_AG_after_on_the_right(&A, A.ie_on_align + 1, &rcut, &rgen, &rest);

  int f_donor_right, f_acceptor_left, f_factor_left;
  if (lgen == 0 && rgen == 0) {             /* already GT-AG */
    f_donor_right = A.new_donor_right_on_gen; f_acceptor_left = A.new_acceptor_left_on_gen; f_factor_left = A.new_acceptor_factor_left;
  } else {
    shiftplan S[4];
    shift_prepare(T, &S[0], &A, est, elen, gen, glen, true, false, "GT");
    shift_prepare(T, &S[1], &A, est, elen, gen, glen, false, false, "GT");
    shift_prepare(T, &S[2], &A, est, elen, gen, glen, true, true, "GC");
    shift_prepare(T, &S[3], &A, est, elen, gen, glen, false, true, "GC");
    dp_wait();
    int dr = 0, al = 0, fl = 0;
    bool done = false;
    for (int k = 0; k < 4 && !done; ++k) {
      int d2 = 0, a2 = 0, f2 = 0;
      if (shift_decide(&S[k], &A, &d2, &a2, &f2)) { dr = d2; al = a2; fl = f2; done = true; }
    }
    if (!done) {
      fl = A.new_acceptor_factor_left; dr = A.new_donor_right_on_gen; al = A.new_acceptor_left_on_gen;
      try_burset(est, elen, gen, glen, &fl, &dr, &al, donor->es, acceptor->ee);
    }
    f_donor_right = dr; f_acceptor_left = al; f_factor_left = fl;
    if (f_acceptor_left > acceptor->ge || f_donor_right < donor->gs) return false;
  }
  donor->ge = f_donor_right;
  acceptor->gs = f_acceptor_left;
  acceptor->es = f_factor_left;
  donor->ee = acceptor->es - 1;
  return true;
}
