// fcd_launch.cuh -- minimal runtime layer under the plan: device memory, copies and the
// phase-kernel launcher.  Two builds of the same sources exist:
//   * default (nvcc, sm_100a): real CUDA; this is the product (libfcd_b200.so).
//   * -DFCD_EMULATE (g++): "device" memory is host memory and a launch runs every block,
//     phase and thread sequentially on the CPU.  TEST INFRASTRUCTURE ONLY (tests/emul); it
//     exists because the build container has no GPU.  The product never loads it.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if !defined(FCD_EMULATE)
#include <cuda_runtime.h>
#endif

namespace fcd {
namespace rt {

using stream_t = void*;

[[noreturn]] inline void fail(const std::string& what) { throw std::runtime_error(what); }
// a failing CUDA runtime call (-> FCD_ERR_RUNTIME); everything else thrown through fail() is a bad argument
struct cuda_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#if defined(FCD_EMULATE)
// ------------------------------------------------------------------ CPU emulation --------
inline void* dmalloc(size_t n) {
    void* p = std::malloc(n ? n : 1);
    if (!p) fail("emul: out of memory");
    return p;
}
inline void dfree(void* p) { std::free(p); }
inline void dmemset(void* p, int v, size_t n, stream_t) { std::memset(p, v, n); }
inline void h2d(void* d, const void* h, size_t n, stream_t) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, stream_t) { std::memcpy(h, d, n); }
inline void d2d(void* d, const void* s, size_t n, stream_t) { std::memmove(d, s, n); }
inline void sync(stream_t) {}
inline int sm_count() { return 148; }

// Order in which the emulated threads of a block run inside one phase: 0 ascending, 1 descending,
// 2 odd threads first.  Within a phase there is no barrier, so a correct kernel must give
// bit-identical results for every order; tests run all three (a poor man's racecheck --
// compute-sanitizer is not available on the GPU pool).
inline int& emu_thread_order() { static int order = 0; return order; }
inline int emu_tid(int i, int n) {
    const int o = emu_thread_order();
    if (o == 1) return n - 1 - i;
    if (o == 2) { const int odd = n / 2; return i < odd ? 2 * i + 1 : 2 * (i - odd); }
    return i;
}

template <class K, int PH>
inline void emu_phases(const typename K::Params& p, int bx, int by, unsigned char* smem, typename K::State* st) {
    std::vector<char> on(K::THREADS);     // evaluated before the phase runs, like the device does
    for (int tid = 0; tid < K::THREADS; ++tid) on[tid] = K::template enabled<PH>(p, smem, tid) ? 1 : 0;
    // a skipped phase also skips its barrier: the decision must be uniform over the barrier
    // domain (whole block, or one group with named barriers) or the GPU deadlocks
    const int dom = K::SYNC_THREADS == 0 ? K::THREADS : K::SYNC_THREADS;
    for (int tid = 0; tid < K::THREADS; ++tid)
        if (on[tid] != on[(tid / dom) * dom]) fail("emul: phase enable flag is not uniform over its barrier domain");
    for (int i = 0; i < K::THREADS; ++i) {
        const int tid = emu_tid(i, K::THREADS);
        if (on[tid]) K::template phase<PH>(p, bx, by, tid, smem, st[tid]);
    }
    if constexpr (PH + 1 < K::PHASES) emu_phases<K, PH + 1>(p, bx, by, smem, st);
}

template <class K>
inline void launch(int gx, int gy, stream_t, const typename K::Params& p) {
    // persistent blocks: a few emulated blocks loop over all tiles, like the device does
    const int ntiles = gx * gy;
    const int grid = ntiles < 3 ? ntiles : 3;
    for (int b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem((size_t)K::SMEM_BYTES + 16, 0xCD);
        std::vector<typename K::State> st(K::THREADS);
        for (int tid = 0; tid < K::THREADS; ++tid) K::prologue(p, tid, smem.data());
        int t0 = b, t1 = ntiles, step = grid;
        if (K::BLOCKED_TILES) {
            t0 = (int)((long long)b * ntiles / grid);
            t1 = (int)((long long)(b + 1) * ntiles / grid);
            step = 1;
        }
        for (int tile = t0; tile < t1; tile += step) {
            if constexpr (K::PIPELINED) {
                const int nt = tile + step;
                for (int tid = 0; tid < K::THREADS; ++tid)
                    st[tid].link = TileLink{nt % gx, nt / gx, nt < t1, tile == t0};
            }
            emu_phases<K, 0>(p, tile % gx, tile / gx, smem.data(), st.data());
        }
    }
}

#else
// ------------------------------------------------------------------ CUDA -----------------
inline void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw cuda_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline int sm_count() {
    int dev = 0, sms = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
    return sms;
}
inline void* dmalloc(size_t n) {
    void* p = nullptr;
    check(cudaMalloc(&p, n ? n : 1), "cudaMalloc");
    return p;
}
inline void dfree(void* p) { if (p) cudaFree(p); }
inline void dmemset(void* p, int v, size_t n, stream_t s) { check(cudaMemsetAsync(p, v, n, (cudaStream_t)s), "cudaMemsetAsync"); }
inline void h2d(void* d, const void* h, size_t n, stream_t s) {
    check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s), "cudaMemcpyAsync h2d");
    check(cudaStreamSynchronize((cudaStream_t)s), "sync after h2d");
}
inline void d2h(void* h, const void* d, size_t n, stream_t s) {
    // Wait for the stream FIRST: a copy into pageable memory that has to wait for queued kernels does so inside the
    // driver, where it blocks the CUDA calls of every other thread of the process (a consumer thread draining a gather
    // ring lost a whole wave per chunk to this); after the wait the copy itself is immediate.
    check(cudaStreamSynchronize((cudaStream_t)s), "sync before d2h");
    check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s), "cudaMemcpyAsync d2h");
    check(cudaStreamSynchronize((cudaStream_t)s), "sync after d2h");
}
inline void d2d(void* d, const void* s_, size_t n, stream_t s) {
    check(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, (cudaStream_t)s), "cudaMemcpyAsync d2d");
}
inline void sync(stream_t s) { check(cudaStreamSynchronize((cudaStream_t)s), "cudaStreamSynchronize"); }

// Barrier between phases.  Kernels whose groups (one transform each, contiguous warps) never
// exchange data use a named barrier per group, so the groups of a block drift apart and
// overlap each other's load / compute / store phases instead of marching in lockstep.
template <class K>
__device__ __forceinline__ void phase_barrier() {
    if constexpr (K::SYNC_THREADS == 0) {
        __syncthreads();
    } else {
        static_assert(K::SYNC_THREADS % 32 == 0 && K::THREADS / K::SYNC_THREADS <= 15, "named barrier per group");
        const int id = 1 + (int)threadIdx.x / K::SYNC_THREADS;
        asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(K::SYNC_THREADS) : "memory");
    }
}

template <class K, int PH>
__device__ __forceinline__ void run_phases(const typename K::Params& p, int bx, int by, unsigned char* smem,
                                           typename K::State& st) {
    // a disabled phase is skipped together with its trailing barrier (uniform per barrier domain)
    if (K::template enabled<PH>(p, smem, (int)threadIdx.x)) {
        K::template phase<PH>(p, bx, by, (int)threadIdx.x, smem, st);
        if constexpr (PH + 1 < K::PHASES) phase_barrier<K>();
    }
    if constexpr (PH + 1 < K::PHASES) run_phases<K, PH + 1>(p, bx, by, smem, st);
}

// Persistent blocks: grid = min(tiles, SMs * resident blocks per SM); each block runs the
// prologue once (twiddle table -> shared memory) and then loops over tiles (bx, by).
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_BLOCKS) fcd_kernel(const __grid_constant__ typename K::Params p, int gx, int ntiles) {
    extern __shared__ __align__(16) unsigned char fcd_smem[];
    typename K::State st;
    K::prologue(p, (int)threadIdx.x, fcd_smem);
    __syncthreads();
    // BLOCKED_TILES: contiguous tile range per block (bx fastest), a block's consecutive tiles
    // share inputs.  Otherwise round robin: tiles that share 128-byte lines (adjacent column /
    // row tiles) run at the same time on neighbouring blocks, so partial lines merge in L2.
    int t0 = blockIdx.x, t1 = ntiles, step = gridDim.x;
    if (K::BLOCKED_TILES) {
        t0 = (int)((long long)blockIdx.x * ntiles / gridDim.x);
        t1 = (int)((long long)(blockIdx.x + 1) * ntiles / gridDim.x);
        step = 1;
    }
    // tile -> (bx, by) = (tile % gx, tile / gx), kept incrementally: two divisions per block instead of two per tile.
    // Only (bx, by) live across a tile, as before: the next tile's coordinates are formed twice (for the prefetch
    // link, which dies early, and at the loop end) rather than carried through a kernel that sits at its register cap.
    if (t0 >= t1) return;
    int bx = t0 % gx, by = t0 / gx;
    int sx = 1, sy = 0;                              // BLOCKED_TILES: step == 1
    if (!K::BLOCKED_TILES) { sx = step % gx; sy = step / gx; }
    for (int tile = t0; tile < t1; tile += step) {
        if constexpr (K::PIPELINED) {
            int nbx = bx + sx, nby = by + sy;
            if (nbx >= gx) { nbx -= gx; ++nby; }
            st.link = TileLink{nbx, nby, tile + step < t1, tile == t0};
        }
        run_phases<K, 0>(p, bx, by, fcd_smem, st);
        phase_barrier<K>();
        bx += sx; by += sy;
        if (bx >= gx) { bx -= gx; ++by; }
    }
}

template <class K>
inline void launch(int gx, int gy, stream_t s, const typename K::Params& p) {
    // Per kernel instantiation AND per device: blocks that fit on the whole device.  Function attributes
    // (the shared-memory opt-in) belong to the device that is current when they are set, and one process
    // may hold plans on several GPUs, so nothing here is cached across devices.
    constexpr int kMaxDevices = 64;
    static std::atomic<int> resident_of[kMaxDevices];
    int dev = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) fail("device ordinal out of range");
    int resident = resident_of[dev].load(std::memory_order_acquire);
    if (!resident) {    // racing first launches compute the same value; the attribute call is idempotent
        if (K::SMEM_BYTES > 48 * 1024)
            check(cudaFuncSetAttribute(fcd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES),
                  "cudaFuncSetAttribute(smem)");
        int sms = 0, per_sm = 0;
        check(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
        check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fcd_kernel<K>, K::THREADS, K::SMEM_BYTES),
              "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        if (per_sm < 1) fail("kernel does not fit on an SM");
        resident = sms * per_sm;
        resident_of[dev].store(resident, std::memory_order_release);
    }
    const int ntiles = gx * gy;
    if (ntiles <= 0) return;
    const int grid = ntiles < resident ? ntiles : resident;
    fcd_kernel<K><<<dim3((unsigned)grid, 1, 1), K::THREADS, K::SMEM_BYTES, (cudaStream_t)s>>>(p, gx, ntiles);
    check(cudaGetLastError(), "kernel launch");
}
#endif

// per-stage device timing with events recorded on the launch stream
struct StageTimer {
    static constexpr int kStages = 7;
#if defined(FCD_EMULATE)
    void reset() {}
    void begin_chunk(stream_t, int) {}
    void mark(stream_t, int) {}
    void collect(double* ms, long long* launches, long long* frames) {
        for (int i = 0; i < kStages; ++i) { ms[i] = 0; launches[i] = 0; frames[i] = 0; }
    }
#else
    struct Rec { cudaEvent_t a, b; int stage, frames; };
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<Rec> recs;
    cudaEvent_t last = nullptr;
    int cur_frames = 0;
    double acc_ms[kStages] = {0};
    long long acc_n[kStages] = {0}, acc_f[kStages] = {0};
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            check(cudaEventCreate(&e), "cudaEventCreate");
            pool.push_back(e);
        }
        return pool[used++];
    }
    void fold() {   // requires the stream to be idle
        for (auto& r : recs) {
            float ms = 0.f;
            check(cudaEventElapsedTime(&ms, r.a, r.b), "cudaEventElapsedTime");
            acc_ms[r.stage] += ms; acc_n[r.stage] += 1; acc_f[r.stage] += r.frames;
        }
        recs.clear();
        used = 0;
    }
    void reset() {
        recs.clear(); used = 0;
        for (int i = 0; i < kStages; ++i) { acc_ms[i] = 0; acc_n[i] = 0; acc_f[i] = 0; }
    }
    void begin_chunk(stream_t s, int frames) {
        if (used + 16 > 60000) { check(cudaStreamSynchronize((cudaStream_t)s), "sync"); fold(); }
        last = get();
        cur_frames = frames;
        check(cudaEventRecord(last, (cudaStream_t)s), "cudaEventRecord");
    }
    void mark(stream_t s, int stage) {
        cudaEvent_t e = get();
        check(cudaEventRecord(e, (cudaStream_t)s), "cudaEventRecord");
        recs.push_back(Rec{last, e, stage, cur_frames});
        last = e;
    }
    void collect(double* ms, long long* launches, long long* frames) {
        check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        fold();
        for (int i = 0; i < kStages; ++i) { ms[i] = acc_ms[i]; launches[i] = acc_n[i]; frames[i] = acc_f[i]; }
    }
    ~StageTimer() { for (auto e : pool) cudaEventDestroy(e); }
#endif
};

// typed device buffer with RAII
template <class T>
struct DevBuf {
    T* ptr = nullptr;
    size_t count = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { dfree(ptr); }
    void alloc(size_t n) {
        if (n == count && ptr) return;
        dfree(ptr);
        ptr = nullptr;
        count = 0;
        ptr = static_cast<T*>(dmalloc(n * sizeof(T)));
        count = n;
    }
    void grow(size_t n) { if (n > count || !ptr) alloc(n); }   // never shrinks
    void upload(const std::vector<T>& h, stream_t s) {
        alloc(h.size());
        h2d(ptr, h.data(), h.size() * sizeof(T), s);
    }
};

}  // namespace rt
}  // namespace fcd
