// Host-side geometry and filter coefficients of the dataset image path (K5): what Pillow computes for
// `img.thumbnail((imsize, imsize), Image.LANCZOS)` (mdir/external/cirtorch/datasets/datahelpers.py:75-82, reached from
// genericdataset.py:86-97). Plain C++ (no CUDA), so tests/ can check it against live Pillow on a machine without a GPU
// through gdt_thumbnail_geometry / gdt_debug_resize_coeffs. Restated from Pillow's published algorithm
// (PIL/Image.py thumbnail / resize, src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc).
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

namespace gdt {

constexpr int kResizePrecisionBits = 32 - 8 - 2;

struct ThumbGeom {
    int out_w, out_h;   // final size (== input size when nothing is to be done)
    int fx, fy;         // integer box-reduction factors applied before the LANCZOS resample (1 = none)
    int resize;         // 0: the image is left alone
};

// Image.thumbnail's aspect-preserving target size + Image.resize's reducing_gap = 2.0 pre-reduction factors
inline ThumbGeom thumbnail_geometry(int w, int h, double imsize) {
    ThumbGeom g{w, h, 1, 1, 0};
    long x = (long)floor(imsize), y = x;
    if (x >= w && y >= h) return g;
    if (x < 1) x = y = 1;      // Pillow would raise further down; callers validate imsize >= 1
    const double aspect = (double)w / (double)h;
    auto pick = [](double number, auto key) -> long {
        const long lo = (long)floor(number), hi = (long)ceil(number);
        const long best = key(lo) <= key(hi) ? lo : hi;      // min(floor, ceil, key=...) keeps the first minimum
        return best > 1 ? best : 1;
    };
    if ((double)x / (double)y >= aspect) {
        const long yy = y;
        x = pick((double)y * aspect, [&](long n) { return fabs(aspect - (double)n / (double)yy); });
    } else {
        const long xx = x;
        y = pick((double)x / aspect, [&](long n) { return n == 0 ? 0.0 : fabs(aspect - (double)xx / (double)n); });
    }
    if (x == w && y == h) return g;
    g.out_w = (int)x;
    g.out_h = (int)y;
    g.resize = 1;
    int fx = (int)((double)w / (double)g.out_w / 2.0), fy = (int)((double)h / (double)g.out_h / 2.0);
    g.fx = fx > 1 ? fx : 1;
    g.fy = fy > 1 ? fy : 1;
    return g;
}

inline double lanczos_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
inline double lanczos_filter(double x) {
    if (-3.0 <= x && x < 3.0) return lanczos_sinc(x) * lanczos_sinc(x / 3);
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc. bounds: (first input index, tap count) per output; kk: [out][ksize].
inline int resize_coeffs(int in_size, float in0, float in1, int out_size, std::vector<int>& bounds, std::vector<int32_t>& kk) {
    double scale, filterscale;
    filterscale = scale = (double)(in1 - in0) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 3.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k((size_t)ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << kResizePrecisionBits))
                                                  : (int)(0.5 + k[x] * (1 << kResizePrecisionBits));
        }
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    return ksize;
}

// Reduce.c division_UINT32(divider, 8): fixed-point reciprocal of the box area
inline uint32_t reduce_multiplier(int area) {
    const uint32_t max_dividend = (1u << 8) * (uint32_t)area;
    const float max_int = (1 << 30) * 4.0f;
    return (uint32_t)(max_int / max_dividend);
}

}  // namespace gdt
