// Developer probe: does this box do NVSwitch multicast (NVLS), and what does an in-switch all-reduce of the gradient bucket cost?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o _ab/nvls_probe tools/probes/nvls_probe.cu -lcuda
//   _ab/nvls_probe [floats] [unroll 1|2|4] [1 = minimum granularity]
//
// ONE process drives every GPU of the box: a multicast object over all devices, one physical allocation per device bound to
// it, and per device a kernel that reduces its slice with multimem.ld_reduce (the switch adds the W copies) and broadcasts the
// mean with multimem.st.  The ranks are synchronised by the host here (the probe times the data phase only).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CU(x)                                                                 \
    do {                                                                      \
        CUresult r_ = (x);                                                    \
        if (r_ != CUDA_SUCCESS) {                                             \
            const char *s_ = nullptr;                                         \
            cuGetErrorString(r_, &s_);                                        \
            printf("{\"nvls\": false, \"failed\": \"%s\", \"error\": \"%s\"}\n", #x, s_ ? s_ : "?"); \
            return 0;                                                         \
        }                                                                     \
    } while (0)
#define RT(x)                                                                 \
    do {                                                                      \
        cudaError_t e_ = (x);                                                 \
        if (e_ != cudaSuccess) {                                              \
            printf("{\"nvls\": false, \"failed\": \"%s\", \"error\": \"%s\"}\n", #x, cudaGetErrorString(e_)); \
            return 0;                                                         \
        }                                                                     \
    } while (0)

__global__ void fill_kernel(float *p, long long n, float base) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = base + (float)(i % 7);
}

template <int U>
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(float *mc, long long q0, long long q1, float inv) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < q1; q += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < q1)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(mc + 4 * (q + u * stride))
                             : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < q1)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * (q + u * stride)), "f"(v[u].x * inv),
                             "f"(v[u].y * inv), "f"(v[u].z * inv), "f"(v[u].w * inv)
                             : "memory");
    }
}

int main(int argc, char **argv) {
    const long long floats = argc > 1 ? atoll(argv[1]) : 2650000;
    const int unroll = argc > 2 ? atoi(argv[2]) : 1;
    const bool min_gran = argc > 3 && atoi(argv[3]) == 1;  // size the object by the MINIMUM multicast granularity instead of the recommended one
    CU(cuInit(0));
    int n = 0;
    CU(cuDeviceGetCount(&n));
    if (n > 8) n = 8;
    std::vector<CUdevice> dev(n);
    for (int d = 0; d < n; ++d) {
        CU(cuDeviceGet(&dev[d], d));
        int ok = 0;
        CU(cuDeviceGetAttribute(&ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]));
        if (!ok) {
            printf("{\"nvls\": false, \"devices\": %d, \"failed\": \"CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED is 0 on device %d\"}\n", n, d);
            return 0;
        }
    }
    if (n < 2) {
        printf("{\"nvls\": false, \"devices\": %d, \"failed\": \"one GPU\"}\n", n);
        return 0;
    }
    for (int d = 0; d < n; ++d) {  // primary contexts
        RT(cudaSetDevice(d));
        RT(cudaFree(0));
    }
    CUmulticastObjectProp mp{};
    mp.numDevices = n;
    mp.handleTypes = 0;
    mp.flags = 0;
    mp.size = (size_t)floats * 4;
    size_t gran = 0;
    CU(cuMulticastGetGranularity(&gran, &mp, min_gran ? CU_MULTICAST_GRANULARITY_MINIMUM : CU_MULTICAST_GRANULARITY_RECOMMENDED));
    const size_t bytes = ((size_t)floats * 4 + gran - 1) / gran * gran;
    mp.size = bytes;
    CUmemGenericAllocationHandle mc;
    CU(cuMulticastCreate(&mc, &mp));
    for (int d = 0; d < n; ++d) CU(cuMulticastAddDevice(mc, dev[d]));
    std::vector<CUmemGenericAllocationHandle> mem(n);
    std::vector<CUdeviceptr> uc(n), mcva(n);
    for (int d = 0; d < n; ++d) {
        RT(cudaSetDevice(d));
        CUmemAllocationProp ap{};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = d;
        CU(cuMemCreate(&mem[d], bytes, &ap, 0));
        CU(cuMulticastBindMem(mc, 0, mem[d], 0, bytes, 0));
        CUmemAccessDesc acc{};
        acc.location = ap.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemAddressReserve(&uc[d], bytes, gran, 0, 0));
        CU(cuMemMap(uc[d], bytes, 0, mem[d], 0));
        CU(cuMemSetAccess(uc[d], bytes, &acc, 1));
        CU(cuMemAddressReserve(&mcva[d], bytes, gran, 0, 0));
        CU(cuMemMap(mcva[d], bytes, 0, mc, 0));
        CU(cuMemSetAccess(mcva[d], bytes, &acc, 1));
    }
    const long long quads = floats / 4, per = (quads + n - 1) / n;
    std::vector<cudaStream_t> st(n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (int d = 0; d < n; ++d) {
        RT(cudaSetDevice(d));
        RT(cudaStreamCreate(&st[d]));
        RT(cudaEventCreate(&e0[d]));
        RT(cudaEventCreate(&e1[d]));
    }
    int sms = 0;
    RT(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float best = 1e30f, first_err = 0.0f;
    for (int it = 0; it < 12; ++it) {
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            fill_kernel<<<1024, 256, 0, st[d]>>>((float *)uc[d], floats, (float)(d + 1));
        }
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            RT(cudaStreamSynchronize(st[d]));
        }
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            const long long q0 = d * per, q1 = q0 + per < quads ? q0 + per : quads;
            const long long want_grid = (per + 256 * unroll - 1) / (256 * unroll);
            const int grid = (int)(want_grid < 4LL * sms ? want_grid : 4LL * sms);
            RT(cudaEventRecord(e0[d], st[d]));
            if (unroll == 1) nvls_allreduce_kernel<1><<<grid, 256, 0, st[d]>>>((float *)mcva[d], q0, q1, 1.0f / n);
            else if (unroll == 2) nvls_allreduce_kernel<2><<<grid, 256, 0, st[d]>>>((float *)mcva[d], q0, q1, 1.0f / n);
            else nvls_allreduce_kernel<4><<<grid, 256, 0, st[d]>>>((float *)mcva[d], q0, q1, 1.0f / n);
            RT(cudaEventRecord(e1[d], st[d]));
        }
        float worst = 0.0f;
        for (int d = 0; d < n; ++d) {
            RT(cudaSetDevice(d));
            RT(cudaStreamSynchronize(st[d]));
            float ms = 0;
            RT(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            worst = ms > worst ? ms : worst;
        }
        if (it >= 2 && worst < best) best = worst;
        if (it == 0) {  // every rank holds the mean of (d + 1 + i % 7) over d = mean(d + 1) + i % 7
            std::vector<float> h(64);
            for (int d = 0; d < n; ++d) {
                RT(cudaSetDevice(d));
                RT(cudaMemcpy(h.data(), (void *)(uc[d] + (size_t)(quads - 16) * 16), 64 * 4, cudaMemcpyDeviceToHost));
                for (int i = 0; i < 64; ++i) {
                    const long long idx = (quads - 16) * 4 + i;
                    const float want = 0.5f * (n + 1) + (float)(idx % 7);
                    const float err = h[i] > want ? h[i] - want : want - h[i];
                    first_err = err > first_err ? err : first_err;
                }
            }
        }
    }
    printf("{\"nvls\": true, \"devices\": %d, \"floats\": %lld, \"granularity\": %zu, \"unroll\": %d, \"us_data_phase_best\": %.2f, \"max_abs_err\": %.3g}\n", n, floats, gran, unroll,
           best * 1e3f, first_err);
    return 0;
}
