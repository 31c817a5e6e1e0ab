/* oracle/mtr_oracle_main.c -- CLI around the oracle with mTR's flags (main.c:48-123).  TEST INFRASTRUCTURE ONLY.
 * Extra flag -s prints the work counters (DP cells etc.) to stderr as one JSON line. */
#define _POSIX_C_SOURCE 200809L
#include "mtr_oracle.h"
#include <stdlib.h>
#include <unistd.h>

int main(int argc, char **argv)
{
    int print_alignment = 0, manhattan = 1, stats = 0, opt;
    float min_ratio = 0.6f;
    while ((opt = getopt(argc, argv, "acm:ps")) != -1) {
        switch (opt) {
        case 'a': print_alignment = 1; break;
        case 'c': break;
        case 'm': min_ratio = atof(optarg); break;
        case 'p': manhattan = 0; break;
        case 's': stats = 1; break;
        default: return EXIT_FAILURE;
        }
    }
    if (optind >= argc) { fprintf(stderr, "The input file name is expected argument after options\n"); return EXIT_FAILURE; }
    mtro_ctx *c = mtro_new(manhattan, min_ratio);
    int n = mtro_process_file(c, argv[optind], print_alignment);
    if (stats) {
        mtro_stats s; mtro_get_stats(c, &s);
        fprintf(stderr, "{\"reads\": %d, \"bases\": %lld, \"dp_calls\": %lld, \"dp_cells\": %lld, \"revise_calls\": %lld, "
                "\"revise_cells\": %lld, \"print_calls\": %lld, \"print_cells\": %lld, \"di_position_passes\": %lld, "
                "\"candidates\": %lld, \"searches\": %lld}\n", n, s.bases, s.dp_calls, s.dp_cells, s.revise_calls,
                s.revise_cells, s.print_calls, s.print_cells, s.di_position_passes, s.candidates, s.searches);
    }
    mtro_free(c);
    return EXIT_SUCCESS;
}
