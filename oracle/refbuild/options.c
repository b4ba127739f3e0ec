/* TEST INFRASTRUCTURE — see options.h in this directory.
 * Hand-written stand-in for the gengetopt-generated parser: defaults from
 * reference src/options.ggo, `--long=value` / `--long value` / short flags on
 * argv, `name = value` lines in the config file (CLI wins: override=0).
 */
#define _GNU_SOURCE
#include "options.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

enum kind { K_STR, K_INT, K_DBL, K_LONG, K_ENUM, K_FLAG };
struct optdesc { const char *name; char shortc; enum kind k; const char *def; };

static const struct optdesc OPTS[] = {
  {"config-file", 'C', K_STR, "config.ini"},
  {"min-factor-length", 'l', K_INT, "15"},
  {"min-intron-length", 'B', K_INT, "40"},
  {"max-intron-length", 0, K_INT, "0"},
  {"min-string-depth-rate", 'd', K_DBL, "0.2"},
  {"max-prefix-discarded-rate", 'p', K_DBL, "0.60"},
  {"max-suffix-discarded-rate", 's', K_DBL, "0.60"},
  {"max-prefix-discarded", 'P', K_INT, "50"},
  {"max-suffix-discarded", 'S', K_INT, "50"},
  {"min-distance-of-splice-sites", 'D', K_INT, "50"},
  {"max-no-of-factorizations", 0, K_INT, "0"},
  {"max-difference-of-coverage", 0, K_DBL, "0.05"},
  {"max-difference-of-no-of-exons", 0, K_INT, "5"},
  {"max-difference-of-gap-length", 0, K_INT, "20"},
  {"complexity-threshold", 0, K_DBL, "20.0"},
  {"retain-externals", 'E', K_ENUM, "true"},
  {"max-pairings-in-CMEG", 0, K_INT, "80"},
  {"max-shortest-pairing-frequence", 0, K_DBL, "0.4"},
  {"suff-pref-length-intron", 0, K_INT, "70"},
  {"suff-pref-length-est", 0, K_INT, "30"},
  {"suff-pref-length-genomic", 0, K_INT, "30"},
  {"no-transitive-reduction", 0, K_FLAG, NULL},
  {"no-short-edge-compaction", 0, K_FLAG, NULL},
  {"max-single-factorization-time", 0, K_LONG, "900"},
};
#define NOPTS ((int)(sizeof(OPTS) / sizeof(OPTS[0])))

struct slot { void *arg; char **orig; unsigned int *given; };

static struct slot slot_of(struct gengetopt_args_info *a, int i) {
#define S(n) { &a->n##_arg, &a->n##_orig, &a->n##_given }
  struct slot t[] = {
    S(config_file), S(min_factor_length), S(min_intron_length), S(max_intron_length),
    S(min_string_depth_rate), S(max_prefix_discarded_rate), S(max_suffix_discarded_rate),
    S(max_prefix_discarded), S(max_suffix_discarded), S(min_distance_of_splice_sites),
    S(max_no_of_factorizations), S(max_difference_of_coverage), S(max_difference_of_no_of_exons),
    S(max_difference_of_gap_length), S(complexity_threshold), S(retain_externals),
    S(max_pairings_in_CMEG), S(max_shortest_pairing_frequence), S(suff_pref_length_intron),
    S(suff_pref_length_est), S(suff_pref_length_genomic),
    { &a->no_transitive_reduction_flag, NULL, &a->no_transitive_reduction_given },
    { &a->no_short_edge_compaction_flag, NULL, &a->no_short_edge_compaction_given },
    S(max_single_factorization_time),
  };
#undef S
  return t[i];
}

static int assign(struct gengetopt_args_info *a, int i, const char *val, int is_default) {
  struct slot s = slot_of(a, i);
  switch (OPTS[i].k) {
  case K_STR:  *(char **)s.arg = strdup(val); break;
  case K_INT:  *(int *)s.arg = (int)strtol(val, NULL, 0); break;
  case K_LONG: *(long *)s.arg = strtol(val, NULL, 0); break;
  case K_DBL:  *(double *)s.arg = strtod(val, NULL); break;
  case K_ENUM:
    if (strcmp(val, "true") == 0) *(enum enum_retain_externals *)s.arg = retain_externals_arg_true;
    else if (strcmp(val, "false") == 0) *(enum enum_retain_externals *)s.arg = retain_externals_arg_false;
    else { fprintf(stderr, "est-fact: invalid argument, \"%s\", for option `--%s'\n", val, OPTS[i].name); return 1; }
    break;
  case K_FLAG: *(int *)s.arg = !*(int *)s.arg; break;
  }
  if (!is_default) {
    *s.given += 1;
    if (s.orig) { free(*s.orig); *s.orig = strdup(val); }
  }
  return 0;
}

static int find_long(const char *name, size_t len) {
  for (int i = 0; i < NOPTS; ++i)
    if (strlen(OPTS[i].name) == len && strncmp(OPTS[i].name, name, len) == 0) return i;
  return -1;
}
static int find_short(char c) {
  for (int i = 0; i < NOPTS; ++i) if (OPTS[i].shortc && OPTS[i].shortc == c) return i;
  return -1;
}

struct cmdline_parser_params *cmdline_parser_params_create(void) {
  struct cmdline_parser_params *p = malloc(sizeof *p);
  p->override = 0; p->initialize = 1; p->check_required = 1; p->check_ambiguity = 0; p->print_errors = 1;
  return p;
}

static void init_defaults(struct gengetopt_args_info *a) {
  memset(a, 0, sizeof *a);
  for (int i = 0; i < NOPTS; ++i) if (OPTS[i].def) assign(a, i, OPTS[i].def, 1);
}

static void print_help(void) {
  printf("Usage: est-fact [OPTIONS]...\nEST factorization Program\n\n");
  for (int i = 0; i < NOPTS; ++i) {
    if (OPTS[i].shortc) printf("  -%c, --%s\n", OPTS[i].shortc, OPTS[i].name);
    else printf("      --%s\n", OPTS[i].name);
  }
}

int cmdline_parser_ext(int argc, char **argv, struct gengetopt_args_info *a, struct cmdline_parser_params *p) {
  if (p->initialize) init_defaults(a);
  for (int k = 1; k < argc; ++k) {
    const char *s = argv[k];
    int i; const char *val = NULL;
    if (strcmp(s, "-h") == 0 || strcmp(s, "--help") == 0 || strcmp(s, "--detailed-help") == 0) { print_help(); exit(0); }
    if (strcmp(s, "-V") == 0 || strcmp(s, "--version") == 0) { printf("est-fact 0.1\n"); exit(0); }
    if (s[0] == '-' && s[1] == '-') {
      const char *eq = strchr(s + 2, '=');
      size_t len = eq ? (size_t)(eq - (s + 2)) : strlen(s + 2);
      i = find_long(s + 2, len);
      if (i < 0) { fprintf(stderr, "est-fact: unrecognized option '%s'\n", s); return 1; }
      if (OPTS[i].k != K_FLAG) {
        if (eq) val = eq + 1;
        else if (k + 1 < argc) val = argv[++k];
        else { fprintf(stderr, "est-fact: option '%s' requires an argument\n", s); return 1; }
      }
    } else if (s[0] == '-' && s[1]) {
      i = find_short(s[1]);
      if (i < 0) { fprintf(stderr, "est-fact: invalid option -- '%c'\n", s[1]); return 1; }
      if (s[2]) val = s + 2;
      else if (k + 1 < argc) val = argv[++k];
      else { fprintf(stderr, "est-fact: option requires an argument -- '%c'\n", s[1]); return 1; }
    } else { fprintf(stderr, "est-fact: unexpected argument '%s'\n", s); return 1; }
    if (*slot_of(a, i).given && !p->override) continue;
    if (assign(a, i, val ? val : "", 0)) return 1;
  }
  return 0;
}

int cmdline_parser_config_file(const char *filename, struct gengetopt_args_info *a, struct cmdline_parser_params *p) {
  FILE *f = fopen(filename, "r");
  if (!f) return 1;
  char line[4096];
  if (p->initialize) init_defaults(a);
  while (fgets(line, sizeof line, f)) {
    char *s = line;
    while (isspace((unsigned char)*s)) ++s;
    if (*s == '#' || *s == '\0') continue;
    char *e = s;
    while (*e && !isspace((unsigned char)*e) && *e != '=') ++e;
    int i = find_long(s, (size_t)(e - s));
    if (i < 0) { fprintf(stderr, "est-fact: unknown option '%.*s' in %s\n", (int)(e - s), s, filename); fclose(f); return 1; }
    while (isspace((unsigned char)*e) || *e == '=') ++e;
    char *v = e;
    size_t n = strlen(v);
    while (n && isspace((unsigned char)v[n - 1])) v[--n] = '\0';
    if (n >= 2 && v[0] == '"' && v[n - 1] == '"') { v[n - 1] = '\0'; ++v; }
    if (*slot_of(a, i).given && !p->override) continue;
    if (assign(a, i, v, 0)) { fclose(f); return 1; }
  }
  fclose(f);
  return 0;
}

int cmdline_parser_required(struct gengetopt_args_info *a, const char *prog_name) {
  (void)a; (void)prog_name;   /* every option of options.ggo is `optional` */
  return 0;
}

int cmdline_parser_file_save(const char *filename, struct gengetopt_args_info *a) {
  FILE *f = fopen(filename, "w");
  if (!f) return 1;
  for (int i = 0; i < NOPTS; ++i) {
    struct slot s = slot_of(a, i);
    if (!*s.given) continue;
    if (OPTS[i].k == K_FLAG) fprintf(f, "%s\n", OPTS[i].name);
    else if (s.orig && *s.orig) fprintf(f, "%s=\"%s\"\n", OPTS[i].name, *s.orig);
    else fprintf(f, "%s\n", OPTS[i].name);
  }
  fclose(f);
  return 0;
}

void cmdline_parser_free(struct gengetopt_args_info *a) {
  for (int i = 0; i < NOPTS; ++i) {
    struct slot s = slot_of(a, i);
    if (s.orig) { free(*s.orig); *s.orig = NULL; }
  }
  free(a->config_file_arg); a->config_file_arg = NULL;
}
