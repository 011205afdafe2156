// Test driver: parses a DIMACS file through this repository's cnf_io API and prints the result as JSON.
#include <cstdio>
#include <vector>
#include "cnf_io/cnf_io.h"
int main(int argc, char **argv)
{
    int v = 0, c = 0, l = 0;
    if (cnf_header_read(argv[1], &v, &c, &l)) { printf("{\"error\": \"header\"}\n"); return 0; }
    std::vector<int> l_c_num(c > 0 ? c : 1, -1), l_val(l > 0 ? l : 1, 0);
    const bool err = cnf_data_read(argv[1], v, c, l, l_c_num.data(), l_val.data());
    printf("{\"v_num\": %d, \"c_num\": %d, \"l_num\": %d, \"error\": %s, \"l_c_num\": [", v, c, l, err ? "true" : "false");
    for (int i = 0; i < c; i++) printf("%s%d", i ? ", " : "", l_c_num[i]);
    printf("], \"l_val\": [");
    for (int i = 0; i < l; i++) printf("%s%d", i ? ", " : "", l_val[i]);
    printf("]}\n");
    return 0;
}
