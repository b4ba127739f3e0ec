/* TEST INFRASTRUCTURE — part of the oracle build recipe, never shipped.
 *
 * Stand-in for the gengetopt-GENERATED options.h of the reference (the
 * reference tree ships only src/options.ggo; `make` runs gengetopt,
 * reference Makefile:579-583, and gengetopt is not installed in this image).
 * It offers exactly the surface src/configuration.c uses
 * (configuration.c:49-170, 252-323): one <opt>_arg/_orig/_given triple per
 * option of options.ggo, the retain-externals enum, the parser-params struct
 * and six functions.  Written from the option list in options.ggo, not from
 * any generated file.
 */
#ifndef ORACLE_OPTIONS_STANDIN_H
#define ORACLE_OPTIONS_STANDIN_H

enum enum_retain_externals { retain_externals__NULL = -1, retain_externals_arg_true = 0, retain_externals_arg_false };

#define GGO_OPT(type, name) type name##_arg; char *name##_orig; unsigned int name##_given

struct gengetopt_args_info {
  GGO_OPT(char *, config_file);
  GGO_OPT(int, min_factor_length);
  GGO_OPT(int, min_intron_length);
  GGO_OPT(int, max_intron_length);
  GGO_OPT(double, min_string_depth_rate);
  GGO_OPT(double, max_prefix_discarded_rate);
  GGO_OPT(double, max_suffix_discarded_rate);
  GGO_OPT(int, max_prefix_discarded);
  GGO_OPT(int, max_suffix_discarded);
  GGO_OPT(int, min_distance_of_splice_sites);
  GGO_OPT(int, max_no_of_factorizations);
  GGO_OPT(double, max_difference_of_coverage);
  GGO_OPT(int, max_difference_of_no_of_exons);
  GGO_OPT(int, max_difference_of_gap_length);
  GGO_OPT(double, complexity_threshold);
  GGO_OPT(enum enum_retain_externals, retain_externals);
  GGO_OPT(int, max_pairings_in_CMEG);
  GGO_OPT(double, max_shortest_pairing_frequence);
  GGO_OPT(int, suff_pref_length_intron);
  GGO_OPT(int, suff_pref_length_est);
  GGO_OPT(int, suff_pref_length_genomic);
  GGO_OPT(long, max_single_factorization_time);
  int no_transitive_reduction_flag;   unsigned int no_transitive_reduction_given;
  int no_short_edge_compaction_flag;  unsigned int no_short_edge_compaction_given;
};

struct cmdline_parser_params {
  int override;
  int initialize;
  int check_required;
  int check_ambiguity;
  int print_errors;
};

struct cmdline_parser_params *cmdline_parser_params_create(void);
int cmdline_parser_ext(int argc, char **argv, struct gengetopt_args_info *a, struct cmdline_parser_params *p);
int cmdline_parser_config_file(const char *filename, struct gengetopt_args_info *a, struct cmdline_parser_params *p);
int cmdline_parser_required(struct gengetopt_args_info *a, const char *prog_name);
int cmdline_parser_file_save(const char *filename, struct gengetopt_args_info *a);
void cmdline_parser_free(struct gengetopt_args_info *a);

#endif
