/*
 * oracle/batch_oracle.cpp -- CPU ORACLE, batch driver.  TEST INFRASTRUCTURE ONLY (see orc_api.h).
 *
 * The reference processes one frame per call on one thread (object_slam/src/main_obj.cpp:424-450:
 * line_lbd_detect::detect_filter_lines, then detect_3d_cuboid::detect_cuboid).  This driver runs that per-frame path of the oracle over
 * a batch, one frame per loop iteration, frames distributed over the host cores with a static OpenMP schedule:
 * the CPU arm of bench.py (`cpu_baseline`, `--impl reference`) and the batch side of the parity tests.  No Python in the loop.
 *   line_mode 0: segments are an input (the detect_cuboid entry point, how orb_object_slam feeds it, Tracking.cc:1583-1590)
 *   line_mode 1: LSD flavour of detect_filter_lines per frame (use_LSD = true, what object_slam sets, main_obj.cpp:365)
 *   line_mode 2: EDLines flavour (use_LSD = false, the class default, line_lbd_allclass.cpp:121)
 *   line_mode 3 / 4: as 1 / 2, but stage (i) is executed by the REFERENCE'S OWN detect_filter_lines (oracle/_ref/liblinelbd_ref.so, compiled
 *                    from /root/reference, looked up at run time next to this library); -3 when that library is not there
 */
#include <dlfcn.h>
#include <malloc.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_api.h"

extern "C" int lsd_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap,
                              float *raw_lines, int cap_raw, int *n_raw_out, double *scaled_out, double *modgrad_out, double *angles_out,
                              int32_t *list_out, int *list_len, int refine_mode);
extern "C" int edl_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap,
                              float *raw_lines, int cap_raw, int *n_raw_out, uint8_t *blur_out, int16_t *dx_out, int16_t *dy_out,
                              int16_t *g_out, uint8_t *dir_out, int32_t *anchors_out, int *n_anchors_out, uint8_t *edge_out);

typedef int (*ref_dfl_fn)(const uint8_t *img, int w, int h, int channels, int use_LSD, float line_length_thres, float *out, int cap);
static ref_dfl_fn ref_detect_filter_lines_sym()
{
    static ref_dfl_fn fn = []() -> ref_dfl_fn {
        Dl_info info;
        if (!dladdr((void *)&lsd_orc_detect, &info) || !info.dli_fname) return nullptr;
        std::string p(info.dli_fname); /* .../oracle/_build/liboracle.so -> .../oracle/_ref/liblinelbd_ref.so */
        const size_t k = p.rfind("/_build/");
        if (k == std::string::npos) return nullptr;
        p = p.substr(0, k) + "/_ref/liblinelbd_ref.so";
        void *h = dlopen(p.c_str(), RTLD_NOW | RTLD_GLOBAL);
        return h ? (ref_dfl_fn)dlsym(h, "ref_detect_filter_lines") : nullptr;
    }();
    return fn;
}
extern "C" int orc_reference_lines_available(void) { return ref_detect_filter_lines_sym() ? 1 : 0; }

extern "C" int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return (int)std::thread::hardware_concurrency();
#endif
}

/* imgs: n_frames x h x stride bytes.  boxes: CSR (box_off n_frames + 1) of [x y w h prob].  lines: CSR (line_off), used when line_mode == 0.
 * out: box_off[n_frames] x topk_cap records, out_counts per box.  Per frame: n_valid, n_cand, n_lines (segments fed to detect_cuboid).
 * Returns 0, or the first non-zero status of a frame. */
extern "C" int orc_detect_frames_batch(const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels, const double *K, const double *Ts,
                                       const double *boxes, const int32_t *box_off, const double *lines, const int32_t *line_off, int line_mode,
                                       float line_length_thres, const orc_params *p, int n_threads, int topk_cap, orc_cuboid *out, int *out_counts,
                                       int64_t *n_valid, int64_t *n_cand, int32_t *n_lines)
{
    int status = 0;
    if (n_threads < 1) n_threads = 1;
    const ref_dfl_fn ref_lines = (line_mode == 3 || line_mode == 4) ? ref_detect_filter_lines_sym() : nullptr;
    if ((line_mode == 3 || line_mode == 4) && !ref_lines) return -3;
    {
        /* every frame allocates a dozen multi-megabyte images; glibc hands such blocks out with mmap / munmap, and on a many-core host the
         * page faults and address-space locking of a hundred threads doing that at once serialise the whole batch.  Keep the blocks in
         * the per-thread malloc arenas instead (a process-wide setting of the test process; set once). */
        static std::once_flag once;
        std::call_once(once, []() {
            mallopt(M_MMAP_THRESHOLD, 1 << 30);
            mallopt(M_TRIM_THRESHOLD, 1 << 30);
            mallopt(M_TOP_PAD, 64 << 20);
        });
    }
    std::mutex mu;
    auto one_frame = [&](int f) {
        const uint8_t *img = imgs + (size_t)f * h * stride;
        std::vector<double> det;
        const double *fl = nullptr;
        int M = 0;
        if (line_mode == 0) {
            fl = lines + (size_t)line_off[f] * 4;
            M = line_off[f + 1] - line_off[f];
        } else {
            const int cap = 8192;
            std::vector<float> seg((size_t)cap * 4);
            int n = ref_lines ? (stride == w * channels ? ref_lines(img, w, h, channels, line_mode == 3, line_length_thres, seg.data(), cap) : -1)
                    : (line_mode == 1) ? lsd_orc_detect(img, w, h, stride, channels, line_length_thres, seg.data(), cap, nullptr, 0, nullptr, nullptr,
                                                      nullptr, nullptr, nullptr, nullptr, 2)
                                     : edl_orc_detect(img, w, h, stride, channels, line_length_thres, seg.data(), cap, nullptr, 0, nullptr, nullptr, nullptr,
                                                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            if (n < 0) n = 0;
            if (n > cap) n = cap;
            det.resize((size_t)n * 4);
            for (size_t i = 0; i < det.size(); i++) det[i] = (double)seg[i]; /* keylines_to_mat is CV_32F; callers convert to MatrixXd (main_obj.cpp:430-433) */
            fl = det.data();
            M = n;
        }
        const int b0 = box_off[f], N = box_off[f + 1] - b0;
        int64_t nc = 0, nv = 0;
        std::vector<orc_cuboid> loc_out;
        std::vector<int> loc_cnt;
        if (!out) loc_out.resize((size_t)(N > 0 ? N : 1) * topk_cap);
        if (!out_counts) loc_cnt.resize(N > 0 ? N : 1);
        const int rc = orc_detect_cuboid(img, w, h, stride, channels, K, Ts + (size_t)f * 16, boxes + (size_t)b0 * 5, N, fl, M, p, topk_cap,
                                         out ? out + (size_t)b0 * topk_cap : loc_out.data(), out_counts ? out_counts + b0 : loc_cnt.data(), &nc, &nv,
                                         nullptr);
        if (rc) {
            std::lock_guard<std::mutex> g(mu);
            if (!status) status = rc;
        }
        if (n_valid) n_valid[f] = nv;
        if (n_cand) n_cand[f] = nc;
        if (n_lines) n_lines[f] = M;
    };
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int f = 0; f < n_frames; f++) one_frame(f);
#else
    /* the same static schedule by hand: thread t takes the contiguous block [t F / T, (t + 1) F / T) */
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
        pool.emplace_back([&, t]() {
            const int lo = (int)((int64_t)n_frames * t / n_threads), hi = (int)((int64_t)n_frames * (t + 1) / n_threads);
            for (int f = lo; f < hi; f++) one_frame(f);
        });
    for (auto &th : pool) th.join();
#endif
    return status;
}

extern "C" int orc_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}
