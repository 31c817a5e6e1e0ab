// placeholder for the host pipeline (handle_one_file / handle_one_read); replaced by pipeline.cpp
#include <cstdio>
#include <cstdlib>
#include "../../include/mtr_b200.h"
extern "C" {
int Manhattan_Distance = 1; float min_match_ratio = 0.6f; int *orgInputString = nullptr;
float time_all, time_memory, time_range, time_period, time_initialize_input_string, time_wrap_around_DP, time_count_table, time_chaining;
int query_counter;
int handle_one_file(char *, int) { fprintf(stderr, "handle_one_file: not built yet\n"); exit(EXIT_FAILURE); }
void handle_one_read(char *, int, int, int) { fprintf(stderr, "handle_one_read: not built yet\n"); exit(EXIT_FAILURE); }
void mtr_flush(void) {}
}
