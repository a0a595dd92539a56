// oracle/_ref recipe (TEST INFRASTRUCTURE): the reference's block partition of a row range over ranks --
// hpat_dist_get_start / _get_end / _get_node_portion / index_rank, static functions of sdc/_distributed.h:77-102 --
// compiled from the header where it lies and exported through forwarding functions (the header declares them static).
#include "sdc/_distributed.h"

extern "C" {
int64_t ref_dist_get_start(int64_t total, int num_pes, int node_id) { return hpat_dist_get_start(total, num_pes, node_id); }
int64_t ref_dist_get_end(int64_t total, int num_pes, int node_id) { return hpat_dist_get_end(total, num_pes, node_id); }
int64_t ref_dist_get_node_portion(int64_t total, int num_pes, int node_id) { return hpat_dist_get_node_portion(total, num_pes, node_id); }
int64_t ref_index_rank(int64_t total, int num_pes, int index) { return index_rank(total, num_pes, index); }
}
