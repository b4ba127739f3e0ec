/* io.c — command line / config.ini, FASTA input and EST preparation.
 *
 * Contract: the 24 options of reference src/options.ggo with the same long/short names and
 * defaults, precedence CLI > config file > defaults (src/configuration.c:252-278), validation ranges of
 * src/configuration.c:49-170, effective values dumped to ./config-dump.ini.  Sequence prep follows
 * src/io-multifasta.c (reader :93-167, header fields :279-504, reverse-complement :506, polyA/T masking :663-828,
 * N tails :830-868).
 */
#define _GNU_SOURCE
#include "ef.h"
#include <ctype.h>
#include <getopt.h>
#include <unistd.h>

enum { K_STR, K_INT, K_DBL, K_LONG, K_BOOLSTR, K_FLAG };
typedef struct optdef { const char *name; int shortc; int kind; const char *def; } optdef;
static const optdef OPTS[] = {
  {"config-file", 'C', K_STR, "config.ini"}, {"min-factor-length", 'l', K_INT, "15"},
  {"min-intron-length", 'B', K_INT, "40"}, {"max-intron-length", 0, K_INT, "0"},
  {"min-string-depth-rate", 'd', K_DBL, "0.2"}, {"max-prefix-discarded-rate", 'p', K_DBL, "0.60"},
  {"max-suffix-discarded-rate", 's', K_DBL, "0.60"}, {"max-prefix-discarded", 'P', K_INT, "50"},
  {"max-suffix-discarded", 'S', K_INT, "50"}, {"min-distance-of-splice-sites", 'D', K_INT, "50"},
  {"max-no-of-factorizations", 0, K_INT, "0"}, {"max-difference-of-coverage", 0, K_DBL, "0.05"},
  {"max-difference-of-no-of-exons", 0, K_INT, "5"}, {"max-difference-of-gap-length", 0, K_INT, "20"},
  {"complexity-threshold", 0, K_DBL, "20.0"}, {"retain-externals", 'E', K_BOOLSTR, "true"},
  {"max-pairings-in-CMEG", 0, K_INT, "80"}, {"max-shortest-pairing-frequence", 0, K_DBL, "0.4"},
  {"suff-pref-length-intron", 0, K_INT, "70"}, {"suff-pref-length-est", 0, K_INT, "30"},
  {"suff-pref-length-genomic", 0, K_INT, "30"}, {"no-transitive-reduction", 0, K_FLAG, NULL},
  {"no-short-edge-compaction", 0, K_FLAG, NULL}, {"max-single-factorization-time", 0, K_LONG, "900"},
};
#define NOPTS ((int)(sizeof OPTS / sizeof OPTS[0]))
/* ours, outside the reference's option set (ignored by config-dump.ini) */
static const char *EXTRA[] = {"threads", "fibers", "devices", "quiet", "no-aux-outputs", "engine"};
#define NEXTRA 6

typedef struct optval { char *s; bool given; } optval;

static void usage(FILE *f) {
  fprintf(f, "Usage: est-fact [OPTIONS]...\nEST factorization Program (B200 build)\n\n"
             "  -h, --help            Print help and exit\n  -V, --version         Print version and exit\n");
  for (int i = 0; i < NOPTS; ++i) {
    if (OPTS[i].shortc) fprintf(f, "  -%c, --%s%s", OPTS[i].shortc, OPTS[i].name, OPTS[i].kind == K_FLAG ? "" : "=VALUE");
    else fprintf(f, "      --%s%s", OPTS[i].name, OPTS[i].kind == K_FLAG ? "" : "=VALUE");
    if (OPTS[i].def) fprintf(f, "  (default=`%s')", OPTS[i].def);
    fputc('\n', f);
  }
  fprintf(f, "\nExecution (this build only):\n      --threads=N  --fibers=N (ESTs in flight per thread-group)  --devices=0,1,..\n"
             "      --quiet  --no-aux-outputs (skip megs.txt, processed-megs.txt, meg-edges.txt)\n"
             "      --engine=auto|daemon|inproc  where the GPU engine runs: the resident server est-factd (started on demand\n"
             "                                   with auto), or inside this process (environment: EST_FACT_ENGINE, EST_FACTD_SOCKET)\n");
}

static int find_opt(const char *name, size_t len) {
  for (int i = 0; i < NOPTS; ++i)
    if (strlen(OPTS[i].name) == len && strncmp(OPTS[i].name, name, len) == 0) return i;
  return -1;
}

static void set_val(optval *v, const char *s) { free(v->s); v->s = strdup(s); v->given = true; }

static void fmt_double(char *buf, size_t n, double d) {   /* the %.10f-then-trim format of configuration.c:204-214 */
  snprintf(buf, n < 14 ? n : 14, "%.10f", d);
  size_t p = strlen(buf);
  while (p > 1 && buf[p - 1] == '0' && buf[p - 2] != '.') buf[--p] = '\0';
}

#define FAIL_IF(cond) do { if (cond) { fprintf(stderr, "* FATAL est-fact: invalid configuration: %s\n", #cond); return 1; } } while (0)

int ef_config_parse(ef_config *c, int argc, char **argv) {
  optval vals[NOPTS];
  memset(vals, 0, sizeof vals);
  memset(c, 0, sizeof *c);
  c->aux_outputs = true;
  { const char *e = getenv("EST_FACT_ENGINE"); snprintf(c->engine, sizeof c->engine, "%s", (e && (!strcmp(e, "daemon") || !strcmp(e, "inproc"))) ? e : "auto"); }
  struct option lo[NOPTS + 16];
  int nlo = 0;
  char shorts[128] = "hV";
  for (int i = 0; i < NOPTS; ++i) {
    lo[nlo++] = (struct option){OPTS[i].name, OPTS[i].kind == K_FLAG ? no_argument : required_argument, NULL, 1000 + i};
    if (OPTS[i].shortc) { size_t q = strlen(shorts); shorts[q] = (char)OPTS[i].shortc; shorts[q + 1] = ':'; shorts[q + 2] = 0; }
  }
  lo[nlo++] = (struct option){"help", no_argument, NULL, 'h'};
  lo[nlo++] = (struct option){"detailed-help", no_argument, NULL, 'h'};
  lo[nlo++] = (struct option){"version", no_argument, NULL, 'V'};
  for (int i = 0; i < NEXTRA; ++i) lo[nlo++] = (struct option){EXTRA[i], (i < 3 || i == 5) ? required_argument : no_argument, NULL, 2000 + i};
  lo[nlo] = (struct option){0, 0, 0, 0};
  optind = 1;
  int ch;
  while ((ch = getopt_long(argc, argv, shorts, lo, NULL)) != -1) {
    if (ch == 'h') { usage(stdout); exit(0); }
    if (ch == 'V') { printf("est-fact 0.1\n"); exit(0); }
    if (ch == '?') { usage(stderr); return 1; }
    if (ch >= 2000) {
      switch (ch - 2000) {
        case 0: c->threads = atoi(optarg); break;
        case 1: c->fibers = atoi(optarg); break;
        case 2: {
          char *s = strdup(optarg), *tok, *sv = NULL;
          for (tok = strtok_r(s, ",", &sv); tok && c->n_devices < 16; tok = strtok_r(NULL, ",", &sv)) c->devices[c->n_devices++] = atoi(tok);
          free(s);
          break;
        }
        case 3: c->quiet = true; break;
        case 4: c->aux_outputs = false; break;
        case 5:
          if (strcmp(optarg, "auto") && strcmp(optarg, "daemon") && strcmp(optarg, "inproc")) { fprintf(stderr, "est-fact: --engine takes auto, daemon or inproc\n"); return 1; }
          snprintf(c->engine, sizeof c->engine, "%s", optarg);
          break;
      }
      continue;
    }
    int i = ch >= 1000 ? ch - 1000 : -1;
    if (i < 0) for (int k = 0; k < NOPTS; ++k) if (OPTS[k].shortc == ch) i = k;
    if (i < 0) { usage(stderr); return 1; }
    set_val(&vals[i], OPTS[i].kind == K_FLAG ? "1" : optarg);
  }
  /* config file: fills only what the command line did not give */
  const char *cfile = vals[0].given ? vals[0].s : OPTS[0].def;
  if (access(cfile, R_OK) == 0) {
    FILE *f = fopen(cfile, "r");
    char line[4096];
    while (f && fgets(line, sizeof line, f)) {
      char *s = line;
      while (isspace((unsigned char)*s)) ++s;
      if (*s == '#' || !*s) continue;
      char *e = s;
      while (*e && !isspace((unsigned char)*e) && *e != '=') ++e;
      int i = find_opt(s, (size_t)(e - s));
      if (i < 0) { fprintf(stderr, "est-fact: unknown option '%.*s' in %s\n", (int)(e - s), s, cfile); fclose(f); return 1; }
      while (isspace((unsigned char)*e) || *e == '=') ++e;
      size_t n = strlen(e);
      while (n && isspace((unsigned char)e[n - 1])) e[--n] = 0;
      if (n >= 2 && e[0] == '"' && e[n - 1] == '"') { e[n - 1] = 0; ++e; }
      if (!vals[i].given) set_val(&vals[i], OPTS[i].kind == K_FLAG ? "1" : e);
    }
    if (f) fclose(f);
  }
#define SV(i) (vals[i].given ? vals[i].s : OPTS[i].def)
  const int min_factor_length = atoi(SV(1));
  FAIL_IF(min_factor_length <= 0);
  c->min_factor_len = (unsigned)min_factor_length;
  c->min_intron_length = atoi(SV(2)); FAIL_IF(c->min_intron_length < 0);
  c->max_intron_length = atoi(SV(3)); FAIL_IF(c->max_intron_length < 0);
  c->min_string_depth_rate = strtod(SV(4), NULL); FAIL_IF(c->min_string_depth_rate < 0.0 || c->min_string_depth_rate > 1.0);
  c->max_prefix_discarded_rate = strtod(SV(5), NULL); FAIL_IF(c->max_prefix_discarded_rate < 0.0 || c->max_prefix_discarded_rate > 1.0);
  c->max_suffix_discarded_rate = strtod(SV(6), NULL); FAIL_IF(c->max_suffix_discarded_rate < 0.0 || c->max_suffix_discarded_rate > 1.0);
  c->max_prefix_discarded = atoi(SV(7)); FAIL_IF(c->max_prefix_discarded < 0);
  c->max_suffix_discarded = atoi(SV(8)); FAIL_IF(c->max_suffix_discarded < 0);
  const int mds = atoi(SV(9)); FAIL_IF(mds < 0); c->max_site_difference = (unsigned)mds;
  c->max_number_of_factorizations = atoi(SV(10)); FAIL_IF(c->max_number_of_factorizations < 0);
  c->max_coverage_diff = strtod(SV(11), NULL); FAIL_IF(c->max_coverage_diff < 0.0 || c->max_coverage_diff > 1.0);
  c->max_exonNUM_diff = atoi(SV(12)); FAIL_IF(c->max_exonNUM_diff < -1);
  c->max_gapLength_diff = atoi(SV(13)); FAIL_IF(c->max_gapLength_diff < -1);
  c->complexity_threshold = strtod(SV(14), NULL); FAIL_IF(c->complexity_threshold <= 0.0);
  if (strcmp(SV(15), "true") != 0 && strcmp(SV(15), "false") != 0) {
    fprintf(stderr, "est-fact: invalid argument, \"%s\", for option `--retain-externals'\n", SV(15));
    return 1;
  }
  c->retain_externals = strcmp(SV(15), "true") == 0;
  const int mp = atoi(SV(16)); FAIL_IF(mp < 0); c->max_pairings_in_MEG = (unsigned)mp;
  c->max_freq_shortest_pairing = strtod(SV(17), NULL); FAIL_IF(c->max_freq_shortest_pairing < 0.0 || c->max_freq_shortest_pairing > 1.0);
  c->suffpref_length_for_intron = atoi(SV(18)); FAIL_IF(c->suffpref_length_for_intron <= 0);
  c->suffpref_length_on_est = atoi(SV(19)); FAIL_IF(c->suffpref_length_on_est <= 0);
  c->suffpref_length_on_gen = atoi(SV(20)); FAIL_IF(c->suffpref_length_on_gen <= 0);
  c->trans_red = !vals[21].given;
  c->short_edge_comp = !vals[22].given;
  const long mt = strtol(SV(23), NULL, 10); FAIL_IF(mt < 0); c->max_single_factorization_time = (unsigned)mt;
  /* config-dump.ini in the order / quoting of gengetopt's file_save: every valued option, then given flags */
  FILE *d = fopen("./config-dump.ini", "w");
  if (d) {
    char b[64];
    for (int i = 0; i < NOPTS; ++i) {
      switch (OPTS[i].kind) {
        case K_STR: fprintf(d, "%s=\"%s\"\n", OPTS[i].name, SV(i)); break;
        case K_INT: fprintf(d, "%s=\"%d\"\n", OPTS[i].name, atoi(SV(i))); break;
        case K_LONG: fprintf(d, "%s=\"%ld\"\n", OPTS[i].name, strtol(SV(i), NULL, 10)); break;
        case K_DBL: fmt_double(b, sizeof b, strtod(SV(i), NULL)); fprintf(d, "%s=\"%s\"\n", OPTS[i].name, b); break;
        case K_BOOLSTR: fprintf(d, "%s=\"%s\"\n", OPTS[i].name, c->retain_externals ? "true" : "false"); break;
        case K_FLAG: if (vals[i].given) fprintf(d, "%s\n", OPTS[i].name); break;
      }
    }
    fclose(d);
  }
  for (int i = 0; i < NOPTS; ++i) free(vals[i].s);
  return 0;
}

/* ---- FASTA ------------------------------------------------------------------------------------------------ */
/* Lines are right-trimmed of bytes < 0x20 (src/util.c:166-173); a record is a '>' line followed by sequence lines
 * concatenated verbatim up to the next '>' line (or a line equal to "#\#", io-multifasta.c:102). */
struct ef_fasta { FILE *f; char *line; size_t lcap; char *pending; };      /* pending: a header line already consumed */

/* Streaming form (the reference reads the whole file into a list first, io-multifasta.c:93-167; a 27 GB mRNA set does not
 * fit that way): one record per call, the file is read through a 4 MB stdio buffer. */
ef_fasta *ef_fasta_open(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return NULL;
  setvbuf(f, NULL, _IOFBF, (size_t)4 << 20);
  ef_fasta *r = calloc(1, sizeof *r);
  r->f = f;
  return r;
}

void ef_fasta_close(ef_fasta *r) { if (!r) return; fclose(r->f); free(r->line); free(r->pending); free(r); }

int ef_fasta_next(ef_fasta *r, ef_seq *out) {     /* 1 = *out holds the next record (id / seq / orig malloc'ed), 0 = end of file */
  ssize_t len;
  ef_buf cur = {0};
  char *hdr = r->pending;
  r->pending = NULL;
  for (;;) {
    len = getline(&r->line, &r->lcap, r->f);
    if (len == -1) break;
    char *line = r->line;
    while (len > 0 && line[len - 1] < ' ') line[--len] = 0;   /* plain (signed) char, as util.c:168 */
    if (line[0] == '>') {
      if (hdr) { r->pending = strdup(line + 1); break; }       /* the next record starts: this one is complete */
      hdr = strdup(line + 1);
    } else if (hdr) {
      if (strcmp(line, "#\\#") == 0) break;                    /* explicit record terminator (io-multifasta.c:102) */
      if (len > 0) buf_write(&cur, line, (size_t)len);
    }
  }
  if (!hdr) { buf_free(&cur); return 0; }
  memset(out, 0, sizeof *out);
  out->id = hdr;
  out->strand = 1;
  out->seq = cur.p ? cur.p : strdup("");
  out->len = (int)cur.len;
  out->orig = strdup(out->seq);
  return 1;
}

int ef_read_fasta(const char *path, ef_seq **out, size_t *n_out) {
  ef_fasta *r = ef_fasta_open(path);
  if (!r) return -1;
  ef_seq *v = NULL, e;
  size_t n = 0, cap = 0;
  while (ef_fasta_next(r, &e)) {
    if (n == cap) { cap = cap ? cap * 2 : 16; v = realloc(v, cap * sizeof *v); }
    v[n++] = e;
  }
  ef_fasta_close(r);
  *out = v; *n_out = n;
  return 0;
}

/* >chrN:absStart:absEnd:+-1 (io-multifasta.c:306-423); anything else falls back to defaults with an ERROR log */
void ef_parse_genomic_header(ef_seq *g) {
  char *h = strdup(g->id), *save = h, *tok[5] = {0};
  int nt = 0;
  char *p = h;
  while (nt < 5) { tok[nt] = strsep(&p, ":"); if (!tok[nt]) break; ++nt; }
  bool ok = nt == 4;
  if (ok) {
    int a = atoi(tok[1]), b = atoi(tok[2]), s = atoi(tok[3]);
    ok = a >= 1 && b >= 1 && (s == 1 || s == -1);
    if (ok) { g->abs_start = a; g->abs_end = b; g->strand = s; snprintf(g->strand_as_read, sizeof g->strand_as_read, "%.10s", tok[3]); }
  }
  if (!ok) {
    fprintf(stderr, "* ERROR The header of the genomic file is not in the correct standard! Guessed values: unknown:1:%d:+1\n", (int)strlen(g->seq));
    g->abs_start = 1; g->abs_end = (int)strlen(g->seq); g->strand = 1; strcpy(g->strand_as_read, "+1");
  }
  free(save);
}

void ef_ntails_removal(ef_seq *g) {
  int pref = 0, n = (int)strlen(g->seq);
  while (g->seq[pref] == 'N') ++pref;
  if (pref) memmove(g->seq, g->seq + pref, (size_t)(n - pref) + 1);
  g->pref_N = pref;
  n -= pref;
  int suff = 0;
  while (suff < n && g->seq[n - 1 - suff] == 'N') ++suff;
  if (suff == n) { fprintf(stderr, "* FATAL The sequence is only composed by Ns.\n"); exit(1); }
  g->seq[n - suff] = 0;
  g->suff_N = suff;
  g->len = n - suff;
}

void ef_set_gb(ef_seq *e) {
  const char *p = strstr(e->id, "/gb=");
  if (!p) p = strstr(e->id, "/GB=");
  if (!p) return;
  p += 4;
  size_t len = 0;
  while (p[len] != ' ' && p[len] != '/' && p[len] != 0) ++len;
  e->gb = strndup(p, len);
}

static char complement(char c) {
  static const char from[] = "AaTtCcGgRrYyMmKkBbVvDdHh", to[] = "TtAaGgCcYyRrKkMmVvBbHhDd";
  const char *q = c ? strchr(from, c) : NULL;
  return q ? to[q - from] : c;
}

void ef_reverse_complement(ef_seq *e) {   /* both seq and orig receive the complement of seq (io-multifasta.c:506-522) */
  int n = (int)strlen(e->seq);
  for (int l = 0, r = n - 1; l <= r; ++l, --r) {
    const char nr = complement(e->seq[l]), nl = complement(e->seq[r]);
    e->seq[r] = nr; e->seq[l] = nl;
    e->orig[r] = nr; e->orig[l] = nl;
  }
}

void ef_set_strand_and_rc(ef_seq *e) {
  const bool refseq = e->gb && e->gb[0] == 'N' && e->gb[1] && e->gb[2] == '_' && (e->gb[1] == 'M' || e->gb[1] == 'R');
  e->strand = 1; e->fixed_strand = false; e->strand_as_read[0] = 0;
  if (refseq) { strcpy(e->strand_as_read, "1"); e->fixed_strand = true; }
  else {
    const char *p = strstr(e->id, "/clone_end=");
    if (!p) p = strstr(e->id, "/CLONE_END=");
    if (p) {
      p += 11;
      int i = 0;
      while (i < 10 && *p && *p != '\'') e->strand_as_read[i++] = *p++;
      e->strand_as_read[i] = 0;
      bool valid = false;
      if (!strcmp(e->strand_as_read, "3")) { e->strand = 1; valid = true; }
      else if (!strcmp(e->strand_as_read, "5")) { e->strand = -1; valid = true; }
      if (valid) {
        p = strstr(e->id, "/fixed_strand=");
        if (!p) p = strstr(e->id, "/FIXED_STRAND=");
        if (p) e->fixed_strand = p[14] == '1';
      }
    }
  }
  if (e->strand == -1) ef_reverse_complement(e);
}

/* polyA / polyT ends -> '*' / '#' (io-multifasta.c:663-828): a 14-wide sliding window must stay >= 72 % A (or T);
 * the masked length is the last A (T) position reached, accepted if the overall fraction up to it is >= 72 %. */
static void mask_end(ef_seq *e, bool suffix) {
  const size_t W = 14, n = strlen(e->seq);
  const double FR = 0.72;
  char *s = e->seq;
#define AT(i) (*(suffix ? &s[n - (i) - 1] : &s[(i)]))
  size_t cA = 0, cT = 0, rA, rT, lastA = 0, lastT = 0, lastAc = 0, lastTc = 0, i;
  for (i = 0; i < W && i < n; ++i) {
    if (AT(i) == 'A') { ++cA; lastA = i; lastAc = cA; }
    if (AT(i) == 'T') { ++cT; lastT = i; lastTc = cT; }
  }
  rA = cA; rT = cT;
  while (i < n && ((double)rA >= FR * (double)W || (double)rT >= FR * (double)W)) {
    if (AT(i - W) == 'A') --rA;
    if (AT(i - W) == 'T') --rT;
    if (AT(i) == 'A') { ++cA; ++rA; lastA = i; lastAc = cA; }
    if (AT(i) == 'T') { ++cT; ++rT; lastT = i; lastTc = cT; }
    ++i;
  }
  if (lastA < W - 1) lastA = W - 1;
  if (lastT < W - 1) lastT = W - 1;
  if ((double)lastAc >= FR * (double)(lastA + 1) || (double)lastTc >= FR * (double)(lastT + 1)) {
    const bool isA = ((double)lastAc) / (double)(lastA + 1) >= ((double)lastTc) / (double)(lastT + 1);
    const size_t mlen = isA ? lastA + 1 : lastT + 1;
    for (i = 0; i < mlen; ++i) AT(i) = isA ? '*' : '#';
    if (suffix) { if (isA) e->suff_polyA = (int)mlen; else e->suff_polyT = (int)mlen; }
    else { if (isA) e->pref_polyA = (int)mlen; else e->pref_polyT = (int)mlen; }
  }
#undef AT
}

void ef_polyAT_substitution(ef_seq *e) {
  e->pref_polyA = e->suff_polyA = e->pref_polyT = e->suff_polyT = -1;
  if (strlen(e->seq) < 14) return;
  mask_end(e, false);
  mask_end(e, true);
}

/* the reverse-complement companion of an EST whose strand is not fixed (main-est-fact.c:70-87, 206-211):
 * copied AFTER masking, so its `orig` is the complement of the masked sequence */
void ef_make_rc_copy(const ef_seq *src, ef_seq *dst) {
  *dst = *src;
  dst->id = strdup(src->id);
  dst->gb = src->gb ? strdup(src->gb) : NULL;
  dst->seq = strdup(src->seq);
  dst->orig = strdup(src->orig);
  ef_reverse_complement(dst);
  dst->strand = -src->strand;
  ef_polyAT_substitution(dst);
}
