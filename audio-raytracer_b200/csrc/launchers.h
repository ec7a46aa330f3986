// launchers.h -- host-callable entry points of the kernel translation units.
#pragma once
#include "fan_dev.cuh"
#include "scene_dev.cuh"

namespace art {

cudaError_t launch_pack(const PackArgs& a, cudaStream_t stream);
size_t trace_smem_bytes(const GeomLayout& L, int nTargets, bool geomInSmem, bool muffleInSmem);
cudaError_t launch_trace(const TraceArgs& a, int numCtas, bool geomInSmem, bool count, cudaStream_t stream);
size_t trace_grid_smem_bytes(const GeomLayout& L, bool geomInSmem);
size_t trace_grid_scratch_bytes(int numCtas);
void trace_grid_plan(int nLocal, int nTargets, int numCtas, int* warps, int* raysPerWarp);
int trace_grid_rotation(int nLocal, int H, int numCtas, int warpsPerCta, bool fullLivedRays, unsigned int* slots);
int perm_grid_rays_per_warp(int nLocal, int nTargets, int numCtas);
cudaError_t launch_trace_grid(const TraceArgs& a, const GridDesc& g, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream);
size_t bounce_smem_bytes(const GeomLayout& L, bool geomInSmem);
cudaError_t launch_bounce(const TraceArgs& a, const GridDesc& g, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream);
size_t query_fan_scratch_bytes(int numCtas);
size_t query_fan_smem_bytes(const GeomLayout& L, bool geomInSmem);
cudaError_t launch_query_fan(const QueryArgs& a, const FanDesc& fans, int numCtas, bool geomInSmem, bool stats, int maxSmemOptin, cudaStream_t stream);
cudaError_t launch_grid_build(const GridBuildArgs& a, cudaStream_t stream);
cudaError_t launch_fan_build(const FanBuildArgs& a, cudaStream_t stream);
size_t fan_build_scratch_bytes(int nFans, int nc);
void fan_build_set_scratch(FanBuildArgs& a, void* scratch);
size_t perm_smem_bytes(const GeomLayout& L, bool geomInSmem);
cudaError_t launch_permeation(const PermArgs& a, int numCtas, bool geomInSmem, int T, cudaStream_t stream);
cudaError_t launch_perm_last(const PermArgs& a, int T, cudaStream_t stream);
size_t perm_grid_smem_bytes(const GeomLayout& L, bool geomInSmem, bool fans);
cudaError_t launch_permeation_grid(const PermArgs& a, const GridDesc& g, const FanDesc* fans, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream);
int perm_binned_slices(int nLocal, int nTargets, int numSms);
size_t perm_binned_cnt_bytes(int nTargets, int slices);
size_t perm_binned_start_bytes(int nTargets);
size_t perm_binned_block_bytes(int nLocal, int nTargets);
size_t perm_binned_smem_bytes(const GeomLayout& L, bool geomInSmem);
cudaError_t launch_permeation_binned(const PermArgs& a, const PermBinArgs& ba, const FanDesc& fans, int numCtas, bool geomInSmem, cudaStream_t stream);
cudaError_t launch_echo_stats(const uint16_t* echo, size_t n, EchoStats* out, bool sequential, int numSms, cudaStream_t stream);
cudaError_t launch_fibonacci(uint16_t* dirs, int n, cudaStream_t stream);
cudaError_t launch_microbench(int kind, int numSms, float* sink, long long* laneOps, cudaStream_t stream);

}  // namespace art
