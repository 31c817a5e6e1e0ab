/* main.c -- the mTR command line on top of libmtr_b200.so.
 * Same flags, messages, exit codes and -c report as /root/reference/main.c:40-123 (getopt "acm:p"); the
 * work itself is handle_one_file() of the library.  Extra knobs come from the environment so that the
 * command line stays a drop-in: MTR_GPUS, MTR_DEVICE, MTR_THREADS, MTR_BATCH_READS, MTR_BATCH_MBASES. */
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include <unistd.h>
#include "../../include/mtr_b200.h"

static void usage(void)
{
    fprintf(stderr, "mTR [-acp] [-m ratio] <fasta file name> \n");
    fprintf(stderr, "-a: Output the alignment between the input sequence and predicted tandem repeat. \n");
    fprintf(stderr, "-c: Print the computation time of each step.\n");
    fprintf(stderr, "-m ratio: Give a minimum match ratio ranging from 0 to 1.\n");
    fprintf(stderr, "-p: Use Pearson's correlation coefficient distance in place of Manhattan distance.\n");
}

int main(int argc, char *argv[])
{
    int print_time = 0, print_alignment = 0, opt;
    min_match_ratio = 0.6f;             /* MIN_MATCH_RATIO, mTR.h:32 */
    Manhattan_Distance = 1;
    while ((opt = getopt(argc, argv, "acm:p")) != -1) {
        switch (opt) {
        case 'a': print_alignment = 1; break;
        case 'c': print_time = 1; break;
        case 'm':
            min_match_ratio = atof(optarg);
            if (0 <= min_match_ratio && min_match_ratio <= 1) break;
            fprintf(stderr, "The input minimum match ratio must range from 0 to 1.\n");
            exit(EXIT_FAILURE);
        case 'p':
            Manhattan_Distance = 0;
            fprintf(stderr, "Pearson's correlation coefficient distance in place of Manhattan distance.\n");
            break;
        default:
            usage();
            exit(EXIT_FAILURE);
        }
    }
    if (optind >= argc) {
        fprintf(stderr, "The input file name is expected argument after options\n");
        exit(EXIT_FAILURE);
    }
    time_all = time_memory = time_range = time_period = time_initialize_input_string = 0;
    time_count_table = time_wrap_around_DP = time_chaining = 0;
    query_counter = 0;
    struct timeval s, e;
    gettimeofday(&s, NULL);
    handle_one_file(argv[optind], print_alignment);
    gettimeofday(&e, NULL);
    time_all = (e.tv_sec - s.tv_sec) + (e.tv_usec - s.tv_usec) * 1.0E-6;
    if (print_time) {
        fprintf(stderr, "Computation time\n");
        fprintf(stderr, "%f\tall\n", time_all);
        fprintf(stderr, "%f\tallocating memory\n", time_memory);
        fprintf(stderr, "%f\tranges\n", time_range);
        fprintf(stderr, "%f\tComputing periods\n", time_period);
        fprintf(stderr, "\t%f\tInitialize the input\n", time_initialize_input_string);
        fprintf(stderr, "\t%f\tcount table generation\n", time_count_table);
        fprintf(stderr, "\t%f\twrap around\n", time_wrap_around_DP);
        fprintf(stderr, "\t%f\tchaining\n", time_chaining);
        fprintf(stderr, "\t%i\tCount of queries\n", query_counter);
    }
    return EXIT_SUCCESS;
}
