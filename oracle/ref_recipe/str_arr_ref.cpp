// oracle/_ref recipe (TEST INFRASTRUCTURE): compiles the reference's OWN string-array natives from where they lie under
// /root/reference -- sdc/native/str_ext/sdc_str_arr.hpp, a header that defines its extern "C" functions inline
// (stable_argsort :565-593, convert_len_arr_to_offset :242-252, is_na :310-314, set_string_array_range :213-240).
// This translation unit is the whole recipe: one include.  No reference source is copied into the repo.
#include "sdc/native/str_ext/sdc_str_arr.hpp"
