// standalone probe: one 2-D TMA box load of bytes, descriptor as __grid_constant__ parameter (variant 0) or in global memory (variant 1)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../cube_slam_b200/csrc/cs_tma.cuh"

template <int BOXW, int BOXH>
__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int variant, int x, int y, uint8_t *out, int *flag, const uint8_t *raw)
{
    __shared__ __align__(128) uint8_t s[BOXW * BOXH];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    if (tid == 0) cs_mbar_init(&bar);
    __syncthreads();
    if (tid == 0) {
        if (variant == 2) { /* plain 1-D bulk copy of the first BOXW * BOXH bytes (no tensor map): checks the mbarrier half alone */
            const uint32_t b = cs_smem_u32(&bar);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(BOXW * BOXH)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cs_smem_u32(s)), "l"(raw), "r"((uint32_t)(BOXW * BOXH)), "r"(b) : "memory");
        } else
            cs_tma_load_2d(variant == 0 ? &tmap : gmap, s, &bar, x, y, BOXW * BOXH);
    }
    const bool ok = cs_mbar_wait(&bar, 0);
    if (!ok && tid == 0) *flag = 1;
    for (int i = tid; i < BOXW * BOXH; i += blockDim.x * blockDim.y) out[i] = s[i];
}

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 640, H = 480;
    std::vector<uint8_t> img(W * H);
    for (int i = 0; i < W * H; i++) img[i] = (uint8_t)((i * 7 + (i / W) * 13) & 255);
    uint8_t *d_img, *d_out;
    int *d_flag;
    cudaMalloc(&d_img, W * H);
    cudaMalloc(&d_out, 48 * 36);
    cudaMalloc(&d_flag, 4);
    cudaMemset(d_flag, 0, 4);
    cudaMemcpy(d_img, img.data(), W * H, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const bool made = cs_make_tmap_bytes(&tm, d_img, W, H, W, 48, 36);
    printf("tensor map made: %d\n", (int)made);
    CUtensorMap *d_tm;
    cudaMalloc(&d_tm, sizeof(tm));
    cudaMemcpy(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    const int x = argc > 2 ? atoi(argv[2]) : 30, y = argc > 3 ? atoi(argv[3]) : 50;
    k_probe<48, 36><<<1, dim3(32, 8)>>>(tm, d_tm, variant, x, y, d_out, d_flag, d_img);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> out(48 * 36);
    int flag = 0;
    cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 36; r++)
        for (int c = 0; c < 48; c++)
            if (out[r * 48 + c] != (variant == 2 ? img[r * 48 + c] : img[(y + r) * W + x + c])) bad++;
    printf("timeout flag %d, mismatches %d\n", flag, bad);
    return 0;
}
