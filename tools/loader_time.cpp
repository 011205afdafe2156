// Times cnf_header_read + cnf_data_read (the cnf_io API, cnf_io.h:13,16) on one DIMACS file.
// Built twice by tools/bench_loader.py: against this repository's loader and against the reference's.
#include <chrono>
#include <cstdio>
#include <vector>
#include "cnf_io/cnf_io.h"
int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const auto t0 = std::chrono::steady_clock::now();
    int v = 0, c = 0, l = 0;
    if (cnf_header_read(argv[1], &v, &c, &l)) return 1;
    std::vector<int> l_c_num(c > 0 ? c : 1), l_val(l > 0 ? l : 1);
    if (cnf_data_read(argv[1], v, c, l, l_c_num.data(), l_val.data())) return 1;
    const auto t1 = std::chrono::steady_clock::now();
    long long sum = 0;
    for (int i = 0; i < l; i++) sum += l_val[i];
    printf("{\"v_num\": %d, \"c_num\": %d, \"l_num\": %d, \"checksum\": %lld, \"ms\": %.3f}\n", v, c, l, sum,
           std::chrono::duration<double, std::milli>(t1 - t0).count());
    return 0;
}
