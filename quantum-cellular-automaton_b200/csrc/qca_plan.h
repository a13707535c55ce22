// Host-side planning interface (internal).
#pragma once
#include <vector>

#include "qca_b200.h"

namespace qca {

// A tile is 2^kTileBits doubles (64 KiB of shared memory): three CTAs per SM.
constexpr int kTileBits = 13;
// Every tile keeps at least 2^kMinLowBits contiguous doubles (128 B) per row so
// global accesses stay full cache lines.
constexpr int kMinLowBits = 4;

// max_cells: 40 for the state-vector engine (index bits), unbounded for chain-level planning
int32_t validate_rule(const qca_rule_t* r, int max_cells = 40);
double spectral_bound(const qca_rule_t& r);
int32_t chebyshev_plan(double z, double tol, std::vector<double>& a);
void plan_passes(int local_bits, std::vector<qca_pass_t>& out);
void plan_passes_v3(int local_bits, int max_cluster_bits, int min_low, std::vector<qca_pass_t>& out);
struct ShardMap;
void plan_shard(const qca_rule_t& r, int world, ShardMap* map, int rank);
int32_t plan_remote(const qca_rule_t& r, int world, int rank, std::vector<qca_remote_op_t>& out);
int32_t plan_rotation(const qca_rule_t& r, int world, int rank, qca_remote_rotation_t* out);

}  // namespace qca
