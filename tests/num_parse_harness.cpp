// Host build of csrc/num_parse.cuh (the same source the CSV kernels compile): one field per input line ->
// "<status> <bits as hex>" per line.  Built by tests/test_num_parse_cpu.py with g++.
//   num_parse_harness f|i < fields.txt
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>

#include "../sdc_b200/csrc/num_parse.cuh"

static const uint64_t kPow5[] = {
#include "../sdc_b200/csrc/pow5_table.inc"
};

int main(int argc, char** argv) {
    const bool as_int = argc > 1 && argv[1][0] == 'i';
    std::string line;
    while (std::getline(std::cin, line)) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(line.data());
        if (as_int) {
            int64_t v = 0;
            const int st = sdcb200::parse_int64_field(s, (int)line.size(), &v);
            printf("%d %lld\n", st, (long long)v);
        } else {
            double d = 0.0;
            const int st = sdcb200::parse_f64_field(s, (int)line.size(), kPow5, &d);
            uint64_t b;
            memcpy(&b, &d, 8);
            printf("%d %016llx\n", st, (unsigned long long)b);
        }
    }
    return 0;
}
