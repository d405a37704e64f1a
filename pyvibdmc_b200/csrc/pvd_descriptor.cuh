// DistIt descriptors on the device: replaces DistIt.run (simulation_utilities/tensorflow_descriptors/
// distance_descriptors.py:177-213) with its helpers atm_atm_dists (:102-113), get_prepped_vec (:154-168),
// dist_matrix (:88-100), sort_atoms (:115-132) and sort_groups (:134-152).  One thread per walker; every operation is a
// correctly rounded IEEE one in NumPy's order (no contraction), so the output equals the reference's bit for bit.
#pragma once
#include "pvd_common.cuh"

constexpr int PVD_DESC_MAX_PAIRS = PVD_MAX_ATOMS * (PVD_MAX_ATOMS - 1) / 2;
enum : int { PVD_DESC_DISTANCE = 0, PVD_DESC_COULOMB = 1, PVD_DESC_SPF = 2 };

struct DistitParams {
    int natoms, method, full_mat, sort;
    int n_atom_lists, ngroups, gsize, pad;
    int atom_lists[PVD_MAX_ATOMS], atom_list_ofs[PVD_MAX_ATOMS + 1];
    int groups[PVD_MAX_ATOMS];
    double pair_scale[PVD_DESC_MAX_PAIRS];            // Coulomb: Z_i Z_j
    double diag[PVD_MAX_ATOMS];                       // Coulomb: 0.5 Z^2.4; otherwise 0
    double r_eq[PVD_MAX_ATOMS * PVD_MAX_ATOMS];       // SPF: per pair (unsorted) or full matrix of the sorted equilibrium structure
};

__device__ __forceinline__ void desc_colnorm(const double *m, int na, double *nrm)
{
    for (int j = 0; j < na; ++j) {
        double acc = 0.0;
        for (int i = 0; i < na; ++i) acc = __dadd_rn(acc, __dmul_rn(m[i * na + j], m[i * na + j]));
        nrm[j] = __dsqrt_rn(acc);
    }
}
// np.sum's order on a contiguous axis (pairwise_sum): left to right below 8 summands; from 8 on, eight interleaved partial
// sums combined as ((0+1)+(2+3))+((4+5)+(6+7)), then the cnt % 8 leftovers
__device__ __forceinline__ double desc_np_sum(const double *v, int cnt)
{
    double acc = 0.0;
    int k = 0;
    if (cnt >= 8) {
        double r[8];
        for (int q = 0; q < 8; ++q) r[q] = v[q];
        for (k = 8; k < cnt - cnt % 8; k += 8)
            for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], v[k + q]);
        acc = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    }
    for (; k < cnt; ++k) acc = __dadd_rn(acc, v[k]);
    return acc;
}
// m <- m[inds][:, inds] (distance_descriptors.py:131,151: the matrix is symmetric)
__device__ __forceinline__ void desc_permute(double *m, double *tmp, int na, const int *inds)
{
    for (int a = 0; a < na; ++a)
        for (int b = 0; b < na; ++b) tmp[a * na + b] = m[inds[a] * na + inds[b]];
    for (int k = 0; k < na * na; ++k) m[k] = tmp[k];
}

__global__ void __launch_bounds__(128) k_distit(const double *__restrict__ xyz, long long n, const DistitParams *__restrict__ P, double *__restrict__ out)
{
    const int na = P->natoms, method = P->method;
    const int npairs = na * (na - 1) / 2;
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n; w += (long long)gridDim.x * blockDim.x) {
        double x[3 * PVD_MAX_ATOMS], m[PVD_MAX_ATOMS * PVD_MAX_ATOMS], tmp[PVD_MAX_ATOMS * PVD_MAX_ATOMS], nrm[PVD_MAX_ATOMS];
        int inds[PVD_MAX_ATOMS];
        for (int k = 0; k < 3 * na; ++k) x[k] = xyz[w * 3 * na + k];
        const bool vec_only = !P->sort && !P->full_mat;
        int p = 0;
        for (int i = 0; i < na; ++i) {
            for (int j = i + 1; j < na; ++j, ++p) {
                const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
                const double r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                const double val = method == PVD_DESC_COULOMB ? __ddiv_rn(P->pair_scale[p], r) : r;
                if (vec_only) out[w * npairs + p] = method == PVD_DESC_SPF ? 1.0 - __ddiv_rn(P->r_eq[p], val) : val;
                else m[i * na + j] = m[j * na + i] = val;
            }
        }
        if (vec_only) continue;
        for (int a = 0; a < na; ++a) m[a * na + a] = P->diag[a];
        if (P->sort) {
            if (P->n_atom_lists > 0) {
                // each sub-list keeps its positions; its members are placed in order of descending column norm
                desc_colnorm(m, na, nrm);
                for (int l = 0; l < P->n_atom_lists; ++l) {
                    const int b0 = P->atom_list_ofs[l], b1 = P->atom_list_ofs[l + 1];
                    for (int u = b0; u < b1; ++u) {
                        const int a = P->atom_lists[u];
                        int rank = 0;
                        for (int t = b0; t < b1; ++t) {
                            const int b = P->atom_lists[t];
                            rank += (nrm[b] > nrm[a] || (nrm[b] == nrm[a] && b < a)) ? 1 : 0;
                        }
                        inds[P->atom_lists[b0 + rank]] = a;
                    }
                }
                desc_permute(m, tmp, na, inds);
            }
            if (P->ngroups > 0) {
                // whole groups trade places in order of the descending sum of their members' column norms (:146).  The
                // reference's fancy-indexed temporary d_mat[:, :, group] lies (member, walker, row) in memory, so this norm
                // reduces a contiguous axis (np.sum's unrolled order from 8 atoms on, unlike the atom sort's)
                for (int j = 0; j < na; ++j) {
                    for (int i = 0; i < na; ++i) tmp[i] = __dmul_rn(m[i * na + j], m[i * na + j]);
                    nrm[j] = __dsqrt_rn(desc_np_sum(tmp, na));
                }
                double tot[PVD_MAX_ATOMS];
                for (int g = 0; g < P->ngroups; ++g) {
                    // a strided axis for the reference (left to right) unless there is a single walker
                    double v[PVD_MAX_ATOMS];
                    for (int k = 0; k < P->gsize; ++k) v[k] = nrm[P->groups[g * P->gsize + k]];
                    double acc = 0.0;
                    if (n == 1) acc = desc_np_sum(v, P->gsize);
                    else for (int k = 0; k < P->gsize; ++k) acc = __dadd_rn(acc, v[k]);
                    tot[g] = acc;
                }
                for (int a = 0; a < na; ++a) inds[a] = a;
                for (int g = 0; g < P->ngroups; ++g) {
                    int rank = 0;
                    for (int h = 0; h < P->ngroups; ++h) rank += (tot[h] > tot[g] || (tot[h] == tot[g] && h < g)) ? 1 : 0;
                    for (int k = 0; k < P->gsize; ++k) inds[P->groups[rank * P->gsize + k]] = P->groups[g * P->gsize + k];
                }
                desc_permute(m, tmp, na, inds);
            }
        }
        const bool spf = method == PVD_DESC_SPF && P->sort;     // an unsorted full matrix is returned as is (:196-197)
        if (P->full_mat) {
            for (int k = 0; k < na * na; ++k) out[w * na * na + k] = spf ? 1.0 - __ddiv_rn(P->r_eq[k], m[k]) : m[k];
        } else {
            p = 0;
            for (int i = 0; i < na; ++i)
                for (int j = i + 1; j < na; ++j, ++p)
                    out[w * npairs + p] = spf ? 1.0 - __ddiv_rn(P->r_eq[i * na + j], m[i * na + j]) : m[i * na + j];
        }
    }
}
