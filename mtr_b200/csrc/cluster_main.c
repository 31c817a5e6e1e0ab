/* cluster_main.c -- mTR_cluster: the optional cross-read clustering post-pass (mtr_cluster_records, cluster.cpp) over the
 * records bin/mTR printed.    mTR file.fa | mTR_cluster [-t threshold] [-m ratio] [-l min_rep_len] [records.tsv]    */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../../include/mtr_b200.h"

int main(int argc, char *argv[])
{
    mtr_cluster_params p = {0.6, 0.3, 0, 1};
    int opt;
    while ((opt = getopt(argc, argv, "t:m:l:")) != -1) {
        switch (opt) {
        case 't': p.mh_distance_threshold = atof(optarg); break;
        case 'm': p.min_match_ratio = atof(optarg); break;
        case 'l': p.min_rep_len = atoi(optarg); break;
        default:
            fprintf(stderr, "mTR_cluster [-t 2-mer distance threshold] [-m min match ratio] [-l min repeat length] [records.tsv]\n");
            return EXIT_FAILURE;
        }
    }
    FILE *fp = stdin;
    if (optind < argc && strcmp(argv[optind], "-") != 0 && !(fp = fopen(argv[optind], "r"))) {
        fprintf(stderr, "fatal error: cannot open %s\n", argv[optind]);
        return EXIT_FAILURE;
    }
    size_t cap = 1 << 20, n = 0;
    char *buf = (char *)malloc(cap);
    for (;;) {
        if (n == cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
        if (!buf) { fprintf(stderr, "fatal error: out of memory\n"); return EXIT_FAILURE; }
        const size_t got = fread(buf + n, 1, cap - n, fp);
        if (got == 0) break;
        n += got;
    }
    char *out = NULL;
    int64_t out_len = 0;
    const int rc = mtr_cluster_records(buf, (int64_t)n, &p, &out, &out_len);
    if (rc) { fprintf(stderr, "mTR_cluster: mtr_cluster_records failed (%d)\n", rc); return EXIT_FAILURE; }
    fwrite(out, 1, (size_t)out_len, stdout);
    mtr_cluster_free(out);
    free(buf);
    return EXIT_SUCCESS;
}
