// Test driver: parses a DIMACS file through this repository's cnf_io API and prints the result as JSON.
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
#include "cnf_io/cnf_io.h"
// With a second argument "csr": the cnf_read_csr extension instead (off, lit in the 2*var+neg encoding).
static int dump_csr(const char *path)
{
    int v = 0, c = 0;
    std::vector<uint64_t> off;
    std::vector<uint32_t> lit;
    const bool err = cnf_read_csr(path, &v, &c, off, lit);
    if (off.empty()) { printf("{\"error\": \"header\"}\n"); return 0; }
    printf("{\"v_num\": %d, \"c_num\": %d, \"error\": %s, \"off\": [", v, c, err ? "true" : "false");
    for (size_t i = 0; i < off.size(); i++) printf("%s%llu", i ? ", " : "", (unsigned long long)off[i]);
    printf("], \"lit\": [");
    for (size_t i = 0; i < lit.size(); i++) printf("%s%u", i ? ", " : "", lit[i]);
    printf("]}\n");
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 2 && std::string(argv[2]) == "csr") return dump_csr(argv[1]);
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
