// Neighbor-matrix construction on the GPU (replaces nvalchemiops.torch.neighbors.neighbor_list as called from
// aimnet/calculators/neighbors.py:106-125 and aimnet/modules/lr.py:388-396).
//
// Canonical contract (oracle/nblist_oracle.py): a pair (i, j, s) is kept iff d2 < rc*rc in float32 with
//     sv_k = ((sx*c0k) + (sy*c1k)) + (sz*c2k);  r_k = (x_j[k] + sv_k) - x_i[k];  d2 = ((rx*rx)+(ry*ry))+(rz*rz)
// every operation individually rounded (no FMA contraction), (j == i, s == 0) excluded, rows sorted by
// (j, sx, sy, sz), unused slots = fill value, zero shifts.
//
// Two builders:
//   * naive  : one warp per centre atom, candidates c = j_local * n_images + image enumerated in canonical order,
//              32 per step, ballot-compacted -> rows are canonical by construction.  Any number of systems,
//              per-system cells, partial pbc.  O(n_sys_atoms * images) per atom.
//   * cells  : single system with a cell: atoms binned on a grid with bins >= rc (or whole-cell images when the
//              cell is thinner than rc), stable radix sort by bin (cub), one warp per centre atom scanning the
//              neighbouring (bin, image) pairs; rows then sorted per row (bitonic in shared memory) when the
//              canonical order is requested.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <vector>

#include "common.cuh"
#include "launchers.cuh"

namespace aimnet {

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float canonical_d2(float xi, float yi, float zi, float xj, float yj, float zj, float sx,
                                              float sy, float sz, const float* __restrict__ c, bool has_cell) {
    float rx, ry, rz;
    if (has_cell) {
        float svx = __fadd_rn(__fadd_rn(__fmul_rn(sx, c[0]), __fmul_rn(sy, c[3])), __fmul_rn(sz, c[6]));
        float svy = __fadd_rn(__fadd_rn(__fmul_rn(sx, c[1]), __fmul_rn(sy, c[4])), __fmul_rn(sz, c[7]));
        float svz = __fadd_rn(__fadd_rn(__fmul_rn(sx, c[2]), __fmul_rn(sy, c[5])), __fmul_rn(sz, c[8]));
        rx = __fsub_rn(__fadd_rn(xj, svx), xi);
        ry = __fsub_rn(__fadd_rn(yj, svy), yi);
        rz = __fsub_rn(__fadd_rn(zj, svz), zi);
    } else {
        rx = __fsub_rn(xj, xi);
        ry = __fsub_rn(yj, yi);
        rz = __fsub_rn(zj, zi);
    }
    return __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
}

__global__ void seg_ptr_kernel(const int32_t* __restrict__ batch_idx, int n, int n_sys, int32_t* __restrict__ ptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (batch_idx == nullptr) {
        if (i == 0) {
            ptr[0] = 0;
            for (int s = 1; s <= n_sys; ++s) ptr[s] = n;
        }
        return;
    }
    int prev = (i == 0) ? -1 : batch_idx[i - 1];
    int cur = (i == n) ? n_sys : batch_idx[i];
    for (int s = prev + 1; s <= cur && s <= n_sys; ++s) ptr[s] = i;
}

// per-system image ranges (oracle/nblist_oracle.py:image_ranges; rule of aimnet/ops.py:171-193)
__global__ void image_range_kernel(const float* __restrict__ cell, const uint8_t* __restrict__ pbc, int n_cells,
                                   float cutoff, int32_t* __restrict__ nimg) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_cells) return;
    const float* c = cell + 9 * s;
    double a[9];
    for (int k = 0; k < 9; ++k) a[k] = (double)c[k];
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    double inv[9];
    inv[0] = (a[4] * a[8] - a[5] * a[7]) / det;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) / det;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) / det;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) / det;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) / det;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) / det;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
    for (int k = 0; k < 3; ++k) {
        double bn = sqrt(inv[k] * inv[k] + inv[3 + k] * inv[3 + k] + inv[6 + k] * inv[6 + k]);  // column k
        int n = (int)ceil((double)cutoff * bn - 1e-9);
        if (n < 1) n = 1;
        bool periodic = (pbc == nullptr) ? true : (pbc[3 * s + k] != 0);
        nimg[3 * s + k] = periodic ? n : 0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// naive builder: one warp per centre atom
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nb_naive_kernel(const float* __restrict__ pos, int n_atoms, float rc2,
                                                       const float* __restrict__ cell, int n_cells,
                                                       const int32_t* __restrict__ nimg,
                                                       const int32_t* __restrict__ batch_idx,
                                                       const int32_t* __restrict__ seg_ptr, int max_nb, int fill,
                                                       int32_t* __restrict__ nbmat, int32_t* __restrict__ shifts,
                                                       int32_t* __restrict__ nnb, int32_t* __restrict__ max_count) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n_atoms) return;
    int i = warp;
    int sys = batch_idx ? batch_idx[i] : 0;
    int s0 = seg_ptr[sys], s1 = seg_ptr[sys + 1];
    bool has_cell = cell != nullptr;
    int csel = has_cell ? (n_cells == 1 ? 0 : sys) : 0;
    const float* c = has_cell ? cell + 9 * csel : nullptr;
    int nx = 0, ny = 0, nz = 0;
    if (has_cell) {
        nx = nimg[3 * csel];
        ny = nimg[3 * csel + 1];
        nz = nimg[3 * csel + 2];
    }
    int wy = 2 * ny + 1, wz = 2 * nz + 1;
    int n_img = (2 * nx + 1) * wy * wz;
    int zero_img = (nx * wy + ny) * wz + nz;
    float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
    long long total = (long long)(s1 - s0) * n_img;
    int count = 0;
    int32_t* row = nbmat + (size_t)i * max_nb;
    int32_t* srow = shifts ? shifts + (size_t)i * max_nb * 3 : nullptr;
    for (long long base = 0; base < total; base += 32) {
        long long cidx = base + lane;
        bool keep = false;
        int j = 0, sx = 0, sy = 0, sz = 0;
        if (cidx < total) {
            int jl = (int)(cidx / n_img);
            int img = (int)(cidx - (long long)jl * n_img);
            j = s0 + jl;
            sx = img / (wy * wz) - nx;
            int rem = img % (wy * wz);
            sy = rem / wz - ny;
            sz = rem % wz - nz;
            if (!(j == i && img == zero_img)) {
                float d2 = canonical_d2(xi, yi, zi, pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], (float)sx, (float)sy,
                                        (float)sz, c, has_cell);
                keep = d2 < rc2;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            int p = count + __popc(m & ((1u << lane) - 1u));
            if (p < max_nb) {
                row[p] = j;
                if (srow) {
                    srow[3 * p] = sx;
                    srow[3 * p + 1] = sy;
                    srow[3 * p + 2] = sz;
                }
            }
        }
        count += __popc(m);
    }
    for (int p = min(count, max_nb) + lane; p < max_nb; p += 32) {
        row[p] = fill;
        if (srow) {
            srow[3 * p] = 0;
            srow[3 * p + 1] = 0;
            srow[3 * p + 2] = 0;
        }
    }
    if (lane == 0) {
        nnb[i] = count;
        atomicMax(max_count, count);
    }
}

// ------------------------------------------------------------------------------------------------------------
// cell-list builder (single system)
// ------------------------------------------------------------------------------------------------------------
struct GridParams {
    float inv[9];       // inverse cell (float32), frac = x @ inv
    float cell[9];
    int nb[3];          // bins per axis
    int reach[3];       // offsets scanned per axis
    int periodic[3];
};

__global__ void bin_atoms_kernel(const float* __restrict__ pos, int n, GridParams gp, uint32_t* __restrict__ keys,
                                 int32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    int b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float f = x * gp.inv[k] + y * gp.inv[3 + k] + z * gp.inv[6 + k];
        int bk = (int)floorf(f * (float)gp.nb[k]);
        bk = max(0, min(gp.nb[k] - 1, bk));
        b[k] = bk;
    }
    keys[i] = (uint32_t)((b[0] * gp.nb[1] + b[1]) * gp.nb[2] + b[2]);
    vals[i] = i;
}

__global__ void bin_start_kernel(const uint32_t* __restrict__ sorted_keys, int n, int n_bins,
                                 int32_t* __restrict__ bin_start) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int prev = (i == 0) ? -1 : (int)sorted_keys[i - 1];
    int cur = (i == n) ? n_bins : (int)sorted_keys[i];
    for (int b = prev + 1; b <= cur; ++b) bin_start[b] = i;
}

// positions in bin order (x, y, z, original index): the scan below then reads a bin's atoms as consecutive 16-byte words
// instead of gathering 12 bytes per lane through the index
__global__ void bin_gather_kernel(const float* __restrict__ pos, int n, const int32_t* __restrict__ sorted_idx,
                                  float4* __restrict__ pos_s) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int j = sorted_idx[p];
    pos_s[p] = make_float4(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], __int_as_float(j));
}

__global__ void __launch_bounds__(256) nb_cells_kernel(const float4* __restrict__ pos_s, int n_atoms, float rc2,
                                                       GridParams gp, const uint32_t* __restrict__ sorted_keys,
                                                       const int32_t* __restrict__ sorted_idx,
                                                       const int32_t* __restrict__ bin_start, int max_nb, int fill,
                                                       int32_t* __restrict__ nbmat, int32_t* __restrict__ shifts,
                                                       int32_t* __restrict__ nnb, int32_t* __restrict__ max_count) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n_atoms) return;
    // process atoms in bin order so that a warp's neighbours are hot in L1/L2
    const float4 own = pos_s[warp];
    int i = __float_as_int(own.w);
    uint32_t key = sorted_keys[warp];
    int bz = key % gp.nb[2];
    int by = (key / gp.nb[2]) % gp.nb[1];
    int bx = key / (gp.nb[2] * gp.nb[1]);
    float xi = own.x, yi = own.y, zi = own.z;
    int count = 0;
    int32_t* row = nbmat + (size_t)i * max_nb;
    int32_t* srow = shifts + (size_t)i * max_nb * 3;
    for (int ox = -gp.reach[0]; ox <= gp.reach[0]; ++ox) {
        int tx = bx + ox, sx = 0;
        if (gp.periodic[0]) {
            sx = (tx >= 0) ? tx / gp.nb[0] : -((-tx + gp.nb[0] - 1) / gp.nb[0]);
            tx -= sx * gp.nb[0];
        } else if (tx < 0 || tx >= gp.nb[0]) continue;
        for (int oy = -gp.reach[1]; oy <= gp.reach[1]; ++oy) {
            int ty = by + oy, sy = 0;
            if (gp.periodic[1]) {
                sy = (ty >= 0) ? ty / gp.nb[1] : -((-ty + gp.nb[1] - 1) / gp.nb[1]);
                ty -= sy * gp.nb[1];
            } else if (ty < 0 || ty >= gp.nb[1]) continue;
            for (int oz = -gp.reach[2]; oz <= gp.reach[2]; ++oz) {
                int tz = bz + oz, sz = 0;
                if (gp.periodic[2]) {
                    sz = (tz >= 0) ? tz / gp.nb[2] : -((-tz + gp.nb[2] - 1) / gp.nb[2]);
                    tz -= sz * gp.nb[2];
                } else if (tz < 0 || tz >= gp.nb[2]) continue;
                int bin = (tx * gp.nb[1] + ty) * gp.nb[2] + tz;
                int p0 = bin_start[bin], p1 = bin_start[bin + 1];
                bool zero = (sx == 0 && sy == 0 && sz == 0);
                for (int base = p0; base < p1; base += 32) {
                    int p = base + lane;
                    bool keep = false;
                    int j = 0;
                    if (p < p1) {
                        const float4 pj = pos_s[p];
                        j = __float_as_int(pj.w);
                        if (!(zero && j == i)) {
                            float d2 = canonical_d2(xi, yi, zi, pj.x, pj.y, pj.z, (float)sx, (float)sy, (float)sz, gp.cell, true);
                            keep = d2 < rc2;
                        }
                    }
                    unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        int q = count + __popc(m & ((1u << lane) - 1u));
                        if (q < max_nb) {
                            row[q] = j;
                            srow[3 * q] = sx;
                            srow[3 * q + 1] = sy;
                            srow[3 * q + 2] = sz;
                        }
                    }
                    count += __popc(m);
                }
            }
        }
    }
    for (int p = min(count, max_nb) + lane; p < max_nb; p += 32) {
        row[p] = fill;
        srow[3 * p] = 0;
        srow[3 * p + 1] = 0;
        srow[3 * p + 2] = 0;
    }
    if (lane == 0) {
        nnb[i] = count;
        atomicMax(max_count, count);
    }
}

// per-row bitonic sort by (j, sx, sy, sz); one block per row, P = power of two >= row length (<= 4096)
template <int P>
__global__ void __launch_bounds__(256) nb_sort_rows_kernel(int n_atoms, int max_nb, const int32_t* __restrict__ nnb,
                                                           int32_t* __restrict__ nbmat, int32_t* __restrict__ shifts) {
    __shared__ unsigned long long keys[P];
    int i = blockIdx.x;
    int cnt = min(nnb[i], max_nb);
    int32_t* row = nbmat + (size_t)i * max_nb;
    int32_t* srow = shifts + (size_t)i * max_nb * 3;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        unsigned long long k = ~0ull;
        if (p < cnt) {
            unsigned long long j = (unsigned)row[p];
            unsigned sx = (unsigned)(srow[3 * p] + 128) & 0xff, sy = (unsigned)(srow[3 * p + 1] + 128) & 0xff,
                     sz = (unsigned)(srow[3 * p + 2] + 128) & 0xff;
            k = (j << 24) | (sx << 16) | (sy << 8) | sz;
        }
        keys[p] = k;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int p = threadIdx.x; p < cnt; p += blockDim.x) {
        unsigned long long k = keys[p];
        row[p] = (int32_t)(k >> 24);
        srow[3 * p] = (int)((k >> 16) & 0xff) - 128;
        srow[3 * p + 1] = (int)((k >> 8) & 0xff) - 128;
        srow[3 * p + 2] = (int)(k & 0xff) - 128;
    }
}

__global__ void wrap_kernel(const float* __restrict__ pos, float* __restrict__ out, int n,
                            const float* __restrict__ cell, int n_cells, const uint8_t* __restrict__ pbc,
                            const int32_t* __restrict__ batch_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = (n_cells == 1 || batch_idx == nullptr) ? 0 : batch_idx[i];
    const float* c = cell + 9 * s;
    float a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = c[k];
    float det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    float id = 1.0f / det;
    float inv[9];
    inv[0] = (a[4] * a[8] - a[5] * a[7]) * id;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) * id;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) * id;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    // The reference computes ((x @ inv(cell)) mod 1) @ cell (aimnet/calculators/neighbors.py:331-381).  The same image is
    // x - floor(x @ inv(cell)) @ cell; written this way an atom that already lies inside the cell keeps its coordinates bit for
    // bit, where the fractional round trip moves it by up to an ulp of the cell length (4e-6 A at 60 A, which stiff
    // potentials turn into 1e-4 eV/A of force noise), and a moved atom is rounded once instead of twice.
    float nsh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = x * inv[k] + y * inv[3 + k] + z * inv[6 + k];
        bool per = pbc == nullptr ? true : (pbc[3 * s + k] != 0);
        // fractional coordinates within a few ulps of the faces count as inside (a wrapped atom re-evaluates to -1e-8 or
        // 1 + 1e-7 as often as not): wrapping is idempotent bit for bit
        nsh[k] = (per && (v < -1.0e-6f || v >= 1.0f + 1.0e-6f)) ? floorf(v) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = pos[3 * i + k];
        p = fmaf(-nsh[0], a[k], p);
        p = fmaf(-nsh[1], a[3 + k], p);
        p = fmaf(-nsh[2], a[6 + k], p);
        out[3 * i + k] = p;
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static bool make_grid(const float* hc, const uint8_t* pbc, float cutoff, int n_atoms, GridParams& gp) {
    double a[9];
    for (int k = 0; k < 9; ++k) a[k] = hc[k];
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (!(std::fabs(det) > 1e-12)) return false;
    double inv[9];
    inv[0] = (a[4] * a[8] - a[5] * a[7]) / det;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) / det;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) / det;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) / det;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) / det;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) / det;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
    double rc = (double)cutoff * 1.0002 + 1e-4;   // safety margin for float32 bin assignment
    long long bins = 1;
    for (int k = 0; k < 3; ++k) {
        double bn = std::sqrt(inv[k] * inv[k] + inv[3 + k] * inv[3 + k] + inv[6 + k] * inv[6 + k]);
        double width = 1.0 / bn;   // perpendicular cell width along axis k
        int nb = (int)std::floor(width / rc);
        if (nb < 1) nb = 1;
        if (nb > 256) nb = 256;
        gp.nb[k] = nb;
        gp.periodic[k] = pbc ? (pbc[k] != 0) : 1;
        double bw = width / nb;
        gp.reach[k] = (int)std::ceil(rc / bw - 1e-12);
        if (gp.reach[k] < 1) gp.reach[k] = 1;
        bins *= nb;
    }
    // do not create vastly more bins than atoms
    while (bins > 8LL * n_atoms + 64) {
        int kmax = 0;
        for (int k = 1; k < 3; ++k)
            if (gp.nb[k] > gp.nb[kmax]) kmax = k;
        if (gp.nb[kmax] <= 1) break;
        bins /= gp.nb[kmax];
        gp.nb[kmax] = (gp.nb[kmax] + 1) / 2;
        bins *= gp.nb[kmax];
        double bn = std::sqrt(inv[kmax] * inv[kmax] + inv[3 + kmax] * inv[3 + kmax] + inv[6 + kmax] * inv[6 + kmax]);
        gp.reach[kmax] = std::max(1, (int)std::ceil(rc / ((1.0 / bn) / gp.nb[kmax]) - 1e-12));
    }
    for (int k = 0; k < 9; ++k) {
        gp.inv[k] = (float)inv[k];
        gp.cell[k] = hc[k];
    }
    return true;
}

int neighbor_matrix_impl(const float* positions, int n_atoms, float cutoff, const float* cell, const float* host_cell,
                         const uint8_t* pbc_host, int n_cells, const int32_t* batch_idx, int n_systems, int max_nb,
                         int fill_value, int sorted, int32_t* nbmat, int32_t* shifts, int32_t* nnb,
                         int* max_count_host, cudaStream_t st, bool prefer_cells, int32_t* scratch,
                         int32_t* pinned_host) {
    // scratch (optional, device): >= n_systems + 3*n_cells + 8 ints owned by the caller, avoids the stream-ordered
    // allocations below; pinned_host (optional): page-locked int for the overflow read-back
    AIM_REQUIRE(n_atoms >= 0 && max_nb >= 1, "neighbor_matrix: bad sizes");
    AIM_REQUIRE(cutoff > 0.f, "neighbor_matrix: cutoff must be positive");
    AIM_REQUIRE((cell == nullptr) == (n_cells == 0), "neighbor_matrix: cell / n_cells mismatch");
    AIM_REQUIRE(cell == nullptr || shifts != nullptr, "neighbor_matrix: shifts output required with a cell");
    AIM_REQUIRE(n_cells == 0 || n_cells == 1 || n_cells == n_systems, "neighbor_matrix: n_cells must be 0, 1 or n_systems");
    if (n_systems < 1) n_systems = 1;
    {
        // The cell path takes its sort buffers from the stream-ordered allocator.  With the default release threshold
        // (0) the pool hands everything back to the driver at the next stream synchronisation -- which this function
        // performs for the overflow read-back -- so every call paid fresh allocations (measured: 1.6 .. 35 ms of
        // jitter in the neighbor phase).  Keep the pool's memory instead.
        static thread_local int pool_dev = -1;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev != pool_dev) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_dev = dev;
        }
    }
    int32_t* d_max = scratch;
    if (!scratch) AIM_CUDA_CHECK(cudaMallocAsync(&d_max, sizeof(int32_t), st));
    AIM_CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int32_t), st));
    float rc2 = cutoff * cutoff;
    if (n_atoms > 0) {
        GridParams gp;
        bool use_cells = prefer_cells && cell != nullptr && host_cell != nullptr && n_systems == 1 && n_cells == 1 &&
                         n_atoms >= 512 && make_grid(host_cell, pbc_host, cutoff, n_atoms, gp);
        if (use_cells) {
            // the scan must not visit the same (bin, image) twice: true by construction (distinct offsets ->
            // distinct (wrapped bin, shift) pairs)
            int n_bins = gp.nb[0] * gp.nb[1] * gp.nb[2];
            uint32_t *keys = nullptr, *keys_s = nullptr;
            int32_t *vals = nullptr, *vals_s = nullptr, *bin_start = nullptr;
            void* tmp = nullptr;
            size_t tmp_bytes = 0;
            AIM_CUDA_CHECK(cudaMallocAsync(&keys, sizeof(uint32_t) * n_atoms * 2, st));
            AIM_CUDA_CHECK(cudaMallocAsync(&vals, sizeof(int32_t) * n_atoms * 2, st));
            AIM_CUDA_CHECK(cudaMallocAsync(&bin_start, sizeof(int32_t) * (n_bins + 1), st));
            float4* pos_s = nullptr;
            AIM_CUDA_CHECK(cudaMallocAsync(&pos_s, sizeof(float4) * n_atoms, st));
            keys_s = keys + n_atoms;
            vals_s = vals + n_atoms;
            int bits = 1;
            while ((1 << bits) < n_bins) ++bits;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_s, vals, vals_s, n_atoms, 0, bits, st);
            AIM_CUDA_CHECK(cudaMallocAsync(&tmp, tmp_bytes, st));
            bin_atoms_kernel<<<(n_atoms + 255) / 256, 256, 0, st>>>(positions, n_atoms, gp, keys, vals);
            AIM_LAUNCH_CHECK();
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_s, vals, vals_s, n_atoms, 0, bits, st);
            g_launch_count += 2;
            bin_start_kernel<<<(n_atoms + 256) / 256, 256, 0, st>>>(keys_s, n_atoms, n_bins, bin_start);
            AIM_LAUNCH_CHECK();
            bin_gather_kernel<<<(n_atoms + 255) / 256, 256, 0, st>>>(positions, n_atoms, vals_s, pos_s);
            AIM_LAUNCH_CHECK();
            nb_cells_kernel<<<(n_atoms + 7) / 8, 256, 0, st>>>(pos_s, n_atoms, rc2, gp, keys_s, vals_s, bin_start,
                                                              max_nb, fill_value, nbmat, shifts, nnb, d_max);
            AIM_LAUNCH_CHECK();
            if (sorted) {
                int P = 32;
                while (P < max_nb) P <<= 1;
                AIM_REQUIRE(P <= 4096, "neighbor_matrix: canonical sort supports rows up to 4096 slots");
#define AIM_SORT_CASE(PP)                                                                                   \
    case PP:                                                                                                \
        nb_sort_rows_kernel<PP><<<n_atoms, (PP >= 512 ? 256 : (PP / 2 < 32 ? 32 : PP / 2)), 0, st>>>(n_atoms, max_nb, nnb, nbmat, shifts); \
        break;
                switch (P) {
                    AIM_SORT_CASE(32)
                    AIM_SORT_CASE(64)
                    AIM_SORT_CASE(128)
                    AIM_SORT_CASE(256)
                    AIM_SORT_CASE(512)
                    AIM_SORT_CASE(1024)
                    AIM_SORT_CASE(2048)
                    AIM_SORT_CASE(4096)
                }
#undef AIM_SORT_CASE
                AIM_LAUNCH_CHECK();
            }
            AIM_CUDA_CHECK(cudaFreeAsync(keys, st));
            AIM_CUDA_CHECK(cudaFreeAsync(vals, st));
            AIM_CUDA_CHECK(cudaFreeAsync(bin_start, st));
            AIM_CUDA_CHECK(cudaFreeAsync(pos_s, st));
            AIM_CUDA_CHECK(cudaFreeAsync(tmp, st));
        } else {
            int32_t *seg = scratch ? scratch + 2 : nullptr, *nimg = nullptr;
            uint8_t* d_pbc = nullptr;
            if (!scratch) AIM_CUDA_CHECK(cudaMallocAsync(&seg, sizeof(int32_t) * (n_systems + 1), st));
            seg_ptr_kernel<<<(n_atoms + 256) / 256, 256, 0, st>>>(batch_idx, n_atoms, n_systems, seg);
            AIM_LAUNCH_CHECK();
            if (cell != nullptr) {
                if (scratch)
                    nimg = scratch + 2 + n_systems + 2;
                else
                    AIM_CUDA_CHECK(cudaMallocAsync(&nimg, sizeof(int32_t) * 3 * n_cells, st));
                if (pbc_host != nullptr) {
                    AIM_CUDA_CHECK(cudaMallocAsync(&d_pbc, 3 * n_cells, st));
                    AIM_CUDA_CHECK(cudaMemcpyAsync(d_pbc, pbc_host, 3 * n_cells, cudaMemcpyHostToDevice, st));
                }
                image_range_kernel<<<(n_cells + 127) / 128, 128, 0, st>>>(cell, d_pbc, n_cells, cutoff, nimg);
                AIM_LAUNCH_CHECK();
            }
            nb_naive_kernel<<<(n_atoms + 7) / 8, 256, 0, st>>>(positions, n_atoms, rc2, cell, n_cells, nimg, batch_idx,
                                                              seg, max_nb, fill_value, nbmat, shifts, nnb, d_max);
            AIM_LAUNCH_CHECK();
            if (!scratch) {
                AIM_CUDA_CHECK(cudaFreeAsync(seg, st));
                if (nimg) AIM_CUDA_CHECK(cudaFreeAsync(nimg, st));
            }
            if (d_pbc) AIM_CUDA_CHECK(cudaFreeAsync(d_pbc, st));
        }
    }
    int rc = AIMNET_OK;
    if (max_count_host != nullptr) {
        int32_t h = 0;
        int32_t* dst = pinned_host ? pinned_host : &h;
        // with the caller's scratch and pinned buffer, slot 1 (the engine's largest-molecule count) rides along
        AIM_CUDA_CHECK(cudaMemcpyAsync(dst, d_max, sizeof(int32_t) * ((scratch && pinned_host) ? 2 : 1), cudaMemcpyDeviceToHost, st));
        AIM_CUDA_CHECK(cudaStreamSynchronize(st));
        h = *dst;
        *max_count_host = h;
        if (h > max_nb) rc = AIMNET_NEIGHBOR_OVERFLOW;
    }
    if (!scratch) AIM_CUDA_CHECK(cudaFreeAsync(d_max, st));
    return rc;
}

int wrap_positions_impl(const float* positions, float* wrapped, int n_atoms, const float* cell, int n_cells,
                        const uint8_t* pbc_host, const int32_t* batch_idx, cudaStream_t st) {
    AIM_REQUIRE(cell != nullptr && n_cells >= 1, "wrap_positions: cell required");
    if (n_atoms == 0) return AIMNET_OK;
    uint8_t* d_pbc = nullptr;
    if (pbc_host != nullptr) {
        AIM_CUDA_CHECK(cudaMallocAsync(&d_pbc, 3 * n_cells, st));
        AIM_CUDA_CHECK(cudaMemcpyAsync(d_pbc, pbc_host, 3 * n_cells, cudaMemcpyHostToDevice, st));
    }
    wrap_kernel<<<(n_atoms + 255) / 256, 256, 0, st>>>(positions, wrapped, n_atoms, cell, n_cells, d_pbc, batch_idx);
    AIM_LAUNCH_CHECK();
    if (d_pbc) AIM_CUDA_CHECK(cudaFreeAsync(d_pbc, st));
    return AIMNET_OK;
}

}  // namespace aimnet

extern "C" int aimnet2_neighbor_matrix(const float* positions, int n_atoms, float cutoff, const float* cell,
                                       const float* host_cell, const uint8_t* pbc, int n_cells,
                                       const int32_t* batch_idx, int n_systems, int max_neighbors, int fill_value,
                                       int sorted, int32_t* nbmat, int32_t* shifts, int32_t* num_neighbors,
                                       int* max_count_host, void* stream) {
    return aimnet::neighbor_matrix_impl(positions, n_atoms, cutoff, cell, host_cell, pbc, n_cells, batch_idx, n_systems,
                                        max_neighbors, fill_value, sorted, nbmat, shifts, num_neighbors, max_count_host,
                                        (cudaStream_t)stream, true);
}

extern "C" int aimnet2_wrap_positions(const float* positions, float* wrapped, int n_atoms, const float* cell, int n_cells,
                                      const uint8_t* pbc_host, const int32_t* batch_idx, void* stream) {
    return aimnet::wrap_positions_impl(positions, wrapped, n_atoms, cell, n_cells, pbc_host, batch_idx,
                                       (cudaStream_t)stream);
}
