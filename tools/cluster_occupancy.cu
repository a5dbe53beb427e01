// How many thread-block clusters of 8 / 16 CTAs (256 threads, given dynamic shared memory) a B200 seats at once:
// the question behind k_score_q_bwd's cluster size (one cluster per joint batch, 8 joint batches in the SGLD step).
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_dummy(double* out) { extern __shared__ double sm[]; if (out) out[0] = sm[0]; }
int main() {
    for (size_t smem : {size_t(48) << 10, size_t(96) << 10, size_t(200) << 10}) {
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cln : {2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
            cudaLaunchAttribute attr;
            cfg.gridDim = dim3(8 * cln); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
            attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = cln; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            int mc = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, k_dummy, &cfg);
            printf("smem %3zu KB  cluster %2d: max active clusters %d (%s)\n", smem >> 10, cln, mc, cudaGetErrorString(e));
        }
    }
    return 0;
}
