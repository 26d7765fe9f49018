// die_cluster.cuh -- the four thread-block-cluster primitives the fused environment step needs (sm_90+):
//
//   cluster_rank()          rank of this CTA in its cluster (%cluster_ctarank)
//   cluster_sync()          barrier.cluster arrive.release + wait.acquire over every thread of the cluster: orders shared
//                           AND global memory accesses of the cluster's CTAs (~380 cycles, B300_MICROARCH.md)
//   cluster_map(p, rank)    generic pointer to the same shared-memory object in CTA `rank` of the cluster (distributed
//                           shared memory: plain loads, stores and atomics work on it; ~215 cycles for a remote CTA)
//   launch_cluster(...)     host side: cudaLaunchKernelEx with cudaLaunchAttributeClusterDimension
//
// Under the CPU emulator (tests/hostsim, DIE_HOSTSIM) the same names are provided by tests/hostsim/cuda_runtime.h: the
// CTAs of a cluster run concurrently as fibers, each with its own shared memory, and cluster_map is pointer arithmetic
// between those buffers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if !defined(DIE_HOSTSIM)
#include <cooperative_groups.h>

namespace die {

__device__ __forceinline__ unsigned cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }

__device__ __forceinline__ void cluster_sync() { cooperative_groups::this_cluster().sync(); }

template <typename T>
__device__ __forceinline__ T* cluster_map(T* p, unsigned rank) {
    return cooperative_groups::this_cluster().map_shared_rank(p, rank);
}

// kern<<<grid, block, smem, stream>>>(arg) with thread-block clusters of `cluster_x` CTAs (grid % cluster_x == 0)
template <typename A>
static inline cudaError_t launch_cluster(void (*kern)(const A), unsigned grid, unsigned block, unsigned cluster_x,
                                         size_t smem, cudaStream_t stream, const A& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

// how many clusters of this shape the device can hold at once (0: the shape cannot be scheduled at all)
template <typename A>
static inline int max_active_clusters(void (*kern)(const A), unsigned block, unsigned cluster_x, size_t smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_x * 64);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // namespace die

#endif  // !DIE_HOSTSIM
