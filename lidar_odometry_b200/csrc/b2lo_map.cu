// b2lo_map.cu — K6/K7: the GPU-resident 2-level Z-order voxel hash and its incremental update.
//
// Replaces map::VoxelMap (/root/reference/src/database/VoxelMap.h:188-332, VoxelMap.cpp):
//   UpdateVoxelMap :128-262 (radius cull :146-158, AddPoint :99-120, RegisterToParent :77-80,
//   UnregisterFromParent :82-97, surfel refit / non-planar purge :187-261), GetSurfelAtPoint :368-386,
//   GetPointCloud :388-403, GetL1Surfels :405-418, Clear :122-126.
//
// The reference is sequential and its containers leak their ORDER into the results (SURVEY.md hard
// part 4): L0 iteration order (= GetPointCloud order, = cull order) and each L1's child-set order
// (= summation order of the surfel fit).  The kernels below reproduce those orders exactly, in parallel:
//   cull    : mark -> scan -> replay of "erase in ascending position order, swap-with-last" on indices only
//             (closed form hole_t <- element n-t when no tail voxel is culled, the usual case; otherwise one
//             thread replays the k erases in shared memory), then the moves are applied in parallel;
//             per affected parent one thread replays its child-set swap-erases in removal order.
//   insert  : per-voxel lists of the new points (atomicExch) sorted by point index -> running mean
//             c = (c*n + p)/(n+1) replayed in input order; new voxels ranked by first-seen index with a
//             prefix scan and appended; children appended to their parent in creation order.
//   surfels : one thread per affected L1: gather children in child-set order, f32 covariance, Jacobi SVD,
//             planarity gate; non-planar parents are purged with all their L0 children.  The purge is an
//             arbitrary-order swap-erase on the dense L0 vector; one thread replays it on indices only
//             (shared memory), then the moves are applied in parallel.
#include <climits>
#include "b2lo_internal.h"

namespace b2 {

enum { US_K = 0, US_S = 1, US_NWORK = 2, US_NNEW = 3, US_NAFF = 4, US_NPURGE = 5, US_KPURGE = 6, US_N0 = 7, US_ERR = 8, US_COUNT = 16 };
enum { CT_N0 = 0, CT_N1 = 1, CT_TOMB0 = 2, CT_TOMB1 = 3, CT_ERR = 4, CT_SURF = 5 };
constexpr int ERR_RANGE = 1, ERR_CAP = 2, ERR_INTERNAL = 4;

// ---- cull ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_cull_mark(MapDev M, int n0, float sx, float sy, float sz, float r2, uint8_t* flag, int* blkcnt) {
  __shared__ int sm[40];
  int ntiles = (n0 + 1023) / 1024;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int pos = tile * 1024 + threadIdx.x;
    int mk = 0;
    if (pos < n0) {
      float4 c = M.l0_cent[pos];
      float d[3] = {c.x - sx, c.y - sy, c.z - sz};
      mk = sqn3(d) > r2;  // (centroid - sensor).squaredNorm() > radius_sq  (VoxelMap.cpp:149-150)
      flag[pos] = (uint8_t)mk;
    }
    int tot;
    block_excl_scan(mk, &tot, sm);
    if (threadIdx.x == 0) blkcnt[tile] = tot;
  }
}
__global__ void __launch_bounds__(1024) k_cull_scan(int n0, const int* blkcnt, int* blkoff, int* us) {
  __shared__ int sm[40];
  int ntiles = (n0 + 1023) / 1024, base = 0;
  for (int t0 = 0; t0 < ntiles; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int c = t < ntiles ? blkcnt[t] : 0, tot;
    int e = block_excl_scan(c, &tot, sm);
    if (t < ntiles) blkoff[t] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) { us[US_K] = base; us[US_S] = n0 - base; us[US_N0] = n0 - base; us[US_NWORK] = 0; }
}
__global__ void __launch_bounds__(1024) k_cull_lists(MapDev M, int n0, const uint8_t* flag, const int* blkoff, int* us, int* removed, int* l1work) {
  __shared__ int sm[40];
  const int k = us[US_K];
  if (k == 0) return;
  int ntiles = (n0 + 1023) / 1024;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int pos = tile * 1024 + threadIdx.x;
    int mk = (pos < n0) ? flag[pos] : 0, tot;
    int pre = blkoff[tile] + block_excl_scan(mk, &tot, sm);
    if (pos >= n0) continue;
    if (mk) {
      removed[pre] = pos;
      int x, y, z;
      morton_key(M.l0_key[pos], x, y, z);
      unsigned long long pk = key_morton(parent_coord(x, M.factor), parent_coord(y, M.factor), parent_coord(z, M.factor));
      int s1 = l1_find(M, pk);
      if (s1 >= 0 && atomicCAS(&M.t1_first[s1], INT_MAX, pre) == INT_MAX) l1work[atomicAdd(&us[US_NWORK], 1)] = s1;
    }
  }
}
// one thread per parent that lost children: replay occupied_children.erase() in removal (= L0 dense) order
__global__ void k_cull_unregister(MapDev M, const uint8_t* flag, int* us, const int* l1work) {
  const int nwork = us[US_NWORK];
  for (int wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nwork; wi += gridDim.x * blockDim.x) {
    int s1 = l1work[wi];
    M.t1_first[s1] = INT_MAX;
    L1Meta* mt = &M.l1_meta[s1];
    unsigned long long k1 = M.l1_tab[s1].key;
    int px, py, pz;
    morton_key(k1 & KEY_MASK, px, py, pz);
    int n = mt->nchild;
    int rpos[27]; uint8_t rcode[27]; int nr = 0;
    for (int ci = 0; ci < n; ++ci) {
      int code = mt->child[ci];
      int cx = px * M.factor + code % 3, cy = py * M.factor + (code / 3) % 3, cz = pz * M.factor + code / 9;
      int s0 = l0_find(M, key_morton(cx, cy, cz));
      if (s0 < 0) continue;
      int pos = (int)M.l0_tab[s0].pos;
      if (flag[pos]) { rpos[nr] = pos; rcode[nr] = (uint8_t)code; ++nr; }
    }
    for (int a = 1; a < nr; ++a) {  // ascending dense position = removal order
      int p = rpos[a]; uint8_t c = rcode[a]; int b = a - 1;
      while (b >= 0 && rpos[b] > p) { rpos[b + 1] = rpos[b]; rcode[b + 1] = rcode[b]; --b; }
      rpos[b + 1] = p; rcode[b + 1] = c;
    }
    for (int a = 0; a < nr; ++a) {
      int idx = 0;
      while (idx < n && mt->child[idx] != rcode[a]) ++idx;
      if (idx < n) { mt->child[idx] = mt->child[n - 1]; --n; }
    }
    mt->nchild = (uint8_t)n;
    if (n < 5) k1 &= ~SURFEL_BIT;  // has_surfel = false, last_child_count kept (VoxelMap.cpp:90-92)
    if (n == 0) { k1 = KEY_TOMB; atomicSub(&M.ctr[CT_N1], 1); atomicAdd(&M.ctr[CT_TOMB1], 1); }
    M.l1_tab[s1].key = k1;
  }
}
__global__ void k_cull_move(MapDev M, int* us, const int* removed, const int* aux) {
  const int k = us[US_K];
  const int s = us[US_S];
  const int* fill = aux + 3 * k + 1;  // fill[t], t = 1..k (k_swap_erase_sim)
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < k; r += gridDim.x * blockDim.x) {
    int pos = removed[r];
    M.l0_tab[M.l0_slot[pos]].key = KEY_TOMB;
    if (pos < s) {
      int src = fill[r + 1];
      uint32_t sl = M.l0_slot[src];
      M.l0_cent[pos] = M.l0_cent[src];
      M.l0_key[pos] = M.l0_key[src];
      M.l0_slot[pos] = sl;
      M.l0_tab[sl].pos = (uint32_t)pos;
    }
    if (r == 0) atomicAdd(&M.ctr[CT_TOMB0], k);
  }
}

// ---- insert ---------------------------------------------------------------------------------------------
__global__ void k_ins_probe(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_m, int* us, int* pslot, int* nxt, FEntry* atab, int alog2) {
  const int m = *d_m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    int x = voxel_coord(p.x, M.voxel), y = voxel_coord(p.y, M.voxel), z = voxel_coord(p.z, M.voxel);
    int ax = voxel_coord(p.x, M.scale1), ay = voxel_coord(p.y, M.scale1), az = voxel_coord(p.z, M.scale1);
    if (!key_in_range(x, y, z) || !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) { atomicOr(&us[US_ERR], ERR_RANGE); pslot[i] = -1; continue; }
    bool ins;
    int s0 = l0_find_or_insert(M, key_morton(x, y, z), &ins);
    if (s0 < 0) { atomicOr(&us[US_ERR], ERR_CAP); pslot[i] = -1; continue; }
    atomicMin(&M.t0_first[s0], i);
    atomicAdd(&M.t0_cnt[s0], 1);
    nxt[i] = atomicExch(&M.t0_head[s0], i);
    pslot[i] = s0;
    // affected_L1.insert(PointToVoxelKey(point, 1))  (VoxelMap.cpp:178-179) — float division by voxel*3
    unsigned long long ak = key_morton(ax, ay, az);
    uint32_t mask = (1u << alog2) - 1u, h = hash_slot(ak, alog2);
    for (;;) {
      unsigned long long kk = *((volatile unsigned long long*)&atab[h].key);
      if (kk == ak) break;
      if (kk == KEY_EMPTY) { unsigned long long old = atomicCAS(&atab[h].key, KEY_EMPTY, ak); if (old == KEY_EMPTY || old == ak) break; }
      h = (h + 1) & mask;
    }
    atomicMin(&atab[h].first, (unsigned)i);
  }
}
// leader (first point of each touched voxel) replays AddPoint over the voxel's points in input order
// weighted = 1: the "points" are voxels of a re-hash (w = point_count), merged as in ApplyTransformAndRehash (VoxelMap.cpp:283-297)
__global__ void k_ins_apply(MapDev M, const float4* __restrict__ pts, const int* __restrict__ d_m, const int* pslot, const int* nxt, int* isnew,
                            float4* newc, int weighted) {
  const int m = *d_m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    int s0 = pslot[i];
    if (s0 < 0 || M.t0_first[s0] != i) { isnew[i] = 0; continue; }
    int cnt = M.t0_cnt[s0];
    uint32_t pos = M.l0_tab[s0].pos;
    float cx, cy, cz; int n;
    if (pos == POS_PENDING) { n = 0; cx = cy = cz = 0.0f; }
    else { float4 c = M.l0_cent[pos]; cx = c.x; cy = c.y; cz = c.z; n = __float_as_int(c.w); }
    auto add = [&](int j) {
      float4 p = pts[j];
      if (weighted) {
        int c2 = __float_as_int(p.w);
        if (n == 0) { cx = p.x; cy = p.y; cz = p.z; n = c2; }
        else {
          float n1 = (float)n, n2 = (float)c2, ns = n1 + n2;
          cx = (cx * n1 + p.x * n2) / ns; cy = (cy * n1 + p.y * n2) / ns; cz = (cz * n1 + p.z * n2) / ns;
          n += c2;
        }
      }
      else if (n == 0) { cx = p.x; cy = p.y; cz = p.z; n = 1; }
      else {
        float fn = (float)n, fn1 = (float)(n + 1);
        cx = (cx * fn + p.x) / fn1; cy = (cy * fn + p.y) / fn1; cz = (cz * fn + p.z) / fn1;
        ++n;
      }
    };
    if (cnt == 1) add(i);
    else if (cnt <= 32) {
      int idx[32]; int q = 0;
      for (int j = M.t0_head[s0]; j >= 0 && q < 32; j = nxt[j]) idx[q++] = j;
      for (int a = 1; a < q; ++a) { int v = idx[a], b = a - 1; while (b >= 0 && idx[b] > v) { idx[b + 1] = idx[b]; --b; } idx[b + 1] = v; }
      for (int a = 0; a < q; ++a) add(idx[a]);
    } else {
      int last = -1;
      for (int a = 0; a < cnt; ++a) {  // selection by repeated list walks (pathological multiplicities only)
        int best = INT_MAX;
        for (int j = M.t0_head[s0]; j >= 0; j = nxt[j]) if (j > last && j < best) best = j;
        add(best); last = best;
      }
    }
    float4 out = make_float4(cx, cy, cz, __int_as_float(n));
    if (pos == POS_PENDING) { newc[i] = out; isnew[i] = 1; }
    else { M.l0_cent[pos] = out; isnew[i] = 0; }
    M.t0_first[s0] = INT_MAX; M.t0_cnt[s0] = 0; M.t0_head[s0] = -1;
  }
}
__global__ void __launch_bounds__(1024) k_ins_scan(MapDev M, const int* __restrict__ d_m, const int* isnew, int* newrank, int* us) {
  __shared__ int sm[40];
  const int m = *d_m;
  int base = 0;
  for (int t0 = 0; t0 < m; t0 += blockDim.x) {
    int i = t0 + threadIdx.x;
    int f = i < m ? isnew[i] : 0, tot;
    int e = block_excl_scan(f, &tot, sm);
    if (i < m) newrank[i] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) {
    us[US_NNEW] = base;
    if ((long long)us[US_N0] + base > (long long)M.l0_cap) atomicOr(&us[US_ERR], ERR_CAP);
  }
}
__global__ void k_ins_place(MapDev M, const int* __restrict__ d_m, int* us, const int* pslot, const int* isnew, const int* newrank, const float4* newc,
                            int* l1slot, int* nxt1) {
  const int m = *d_m;
  if (us[US_ERR] & ERR_CAP) return;
  const int base = us[US_N0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    if (!isnew[i]) continue;
    int s0 = pslot[i];
    int pos = base + newrank[i];
    unsigned long long key = M.l0_tab[s0].key;
    M.l0_cent[pos] = newc[i];
    M.l0_key[pos] = key;
    M.l0_slot[pos] = (uint32_t)s0;
    M.l0_tab[s0].pos = (uint32_t)pos;
    int x, y, z;
    morton_key(key, x, y, z);
    bool ins;
    int s1 = l1_find_or_insert(M, key_morton(parent_coord(x, M.factor), parent_coord(y, M.factor), parent_coord(z, M.factor)), &ins);
    l1slot[i] = s1;
    if (s1 < 0) { atomicOr(&us[US_ERR], ERR_CAP); continue; }
    if (ins) {
      L1Meta* mt = &M.l1_meta[s1];
      mt->nchild = 0; mt->planarity = 1.0f; mt->last_child_count = 0;
      M.l1_tab[s1].n[0] = 0.0f; M.l1_tab[s1].n[1] = 0.0f; M.l1_tab[s1].n[2] = 0.0f;
      M.l1_tab[s1].c[0] = 0.0f; M.l1_tab[s1].c[1] = 0.0f; M.l1_tab[s1].c[2] = 0.0f;
      atomicAdd(&M.ctr[CT_N1], 1);
    }
    atomicMin(&M.t1_first[s1], newrank[i]);
    nxt1[i] = atomicExch(&M.t1_head[s1], i);
  }
}
// per parent: append the new children in creation order (RegisterToParent, VoxelMap.cpp:77-80)
__global__ void k_reg_children(MapDev M, const int* __restrict__ d_m, int* us, const int* pslot, const int* isnew, const int* newrank, const int* l1slot,
                               const int* nxt1) {
  const int m = *d_m;
  if (us[US_ERR] & ERR_CAP) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    if (!isnew[i]) continue;
    int s1 = l1slot[i];
    if (s1 < 0 || M.t1_first[s1] != newrank[i]) continue;
    int px, py, pz;
    morton_key(M.l1_tab[s1].key & KEY_MASK, px, py, pz);
    int rk[27]; uint8_t code[27]; int q = 0;
    for (int j = M.t1_head[s1]; j >= 0 && q < 27; j = nxt1[j]) {
      int x, y, z;
      morton_key(M.l0_tab[pslot[j]].key, x, y, z);
      rk[q] = newrank[j];
      code[q] = (uint8_t)((x - px * M.factor) + 3 * (y - py * M.factor) + 9 * (z - pz * M.factor));
      ++q;
    }
    for (int a = 1; a < q; ++a) {
      int v = rk[a]; uint8_t c = code[a]; int b = a - 1;
      while (b >= 0 && rk[b] > v) { rk[b + 1] = rk[b]; code[b + 1] = code[b]; --b; }
      rk[b + 1] = v; code[b + 1] = c;
    }
    L1Meta* mt = &M.l1_meta[s1];
    int n = mt->nchild;
    for (int a = 0; a < q && n < 27; ++a) mt->child[n++] = code[a];
    mt->nchild = (uint8_t)n;
    M.t1_first[s1] = INT_MAX; M.t1_head[s1] = -1;
  }
}

// ---- surfels ----------------------------------------------------------------------------------------------
__global__ void k_surfel(MapDev M, const FEntry* __restrict__ atab, int alog2, int* us, int* plist, unsigned int* pfirst) {
  if (us[US_ERR] & ERR_CAP) return;
  const int acap = 1 << alog2;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < acap; a += gridDim.x * blockDim.x) {
    unsigned long long ak = atab[a].key;
    if (ak == KEY_EMPTY) continue;
    int s1 = l1_find(M, ak);
    if (s1 < 0) continue;
    L1Meta* mt = &M.l1_meta[s1];
    unsigned long long k1 = M.l1_tab[s1].key;
    int N = mt->nchild;
    if (N < 5) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    if ((k1 & SURFEL_BIT) && mt->last_child_count == N) continue;  // incremental skip (VoxelMap.cpp:202-205)
    int px, py, pz;
    morton_key(k1 & KEY_MASK, px, py, pz);
    float cents[27 * 3]; int nc = 0;
    for (int ci = 0; ci < N; ++ci) {
      int code = mt->child[ci];
      int s0 = l0_find(M, key_morton(px * M.factor + code % 3, py * M.factor + (code / 3) % 3, pz * M.factor + code / 9));
      if (s0 < 0) continue;
      float4 c = M.l0_cent[M.l0_tab[s0].pos];
      cents[nc * 3] = c.x; cents[nc * 3 + 1] = c.y; cents[nc * 3 + 2] = c.z; ++nc;
    }
    if (nc < 3) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    float mu[3], nrm[3], plan;
    fit_plane(cents, nc, mu, nrm, &plan);
    if (plan > M.planarity_thr) {  // not planar: the parent and all its children go (VoxelMap.cpp:244-253)
      int idx = atomicAdd(&us[US_NPURGE], 1);
      plist[idx] = s1; pfirst[idx] = atab[a].first;
      continue;
    }
    L1Entry* e = &M.l1_tab[s1];
    e->n[0] = nrm[0]; e->n[1] = nrm[1]; e->n[2] = nrm[2];
    e->c[0] = mu[0]; e->c[1] = mu[1]; e->c[2] = mu[2];
    mt->planarity = plan; mt->last_child_count = N;
    __threadfence();
    e->key = k1 | SURFEL_BIT;
  }
}

// RecomputeAllSurfels (VoxelMap.cpp:304-366): every L1; non-planar parents only lose the surfel (no purge)
__global__ void k_surfel_all(MapDev M) {
  const int tcap = 1 << M.l1_log2cap;
  for (int s1 = blockIdx.x * blockDim.x + threadIdx.x; s1 < tcap; s1 += gridDim.x * blockDim.x) {
    unsigned long long k1 = M.l1_tab[s1].key;
    if (k1 == KEY_EMPTY || k1 == KEY_TOMB) continue;
    L1Meta* mt = &M.l1_meta[s1];
    int N = mt->nchild;
    if (N < 5) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    int px, py, pz;
    morton_key(k1 & KEY_MASK, px, py, pz);
    float cents[27 * 3]; int nc = 0;
    for (int ci = 0; ci < N; ++ci) {
      int code = mt->child[ci];
      int s0 = l0_find(M, key_morton(px * M.factor + code % 3, py * M.factor + (code / 3) % 3, pz * M.factor + code / 9));
      if (s0 < 0) continue;
      float4 c = M.l0_cent[M.l0_tab[s0].pos];
      cents[nc * 3] = c.x; cents[nc * 3 + 1] = c.y; cents[nc * 3 + 2] = c.z; ++nc;
    }
    if (nc < 5) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    float mu[3], nrm[3], plan;
    fit_plane(cents, nc, mu, nrm, &plan);
    if (plan > M.planarity_thr) { M.l1_tab[s1].key = k1 & ~SURFEL_BIT; continue; }
    L1Entry* e = &M.l1_tab[s1];
    e->n[0] = nrm[0]; e->n[1] = nrm[1]; e->n[2] = nrm[2];
    e->c[0] = mu[0]; e->c[1] = mu[1]; e->c[2] = mu[2];
    mt->planarity = plan; mt->last_child_count = N;
    e->key = k1 | SURFEL_BIT;
  }
}
// new_centroid = R * centroid + t (VoxelMap.cpp:273-276), point_count carried in w
struct Rt12 { float R[9]; float t[3]; };
__global__ void k_xform_l0(MapDev M, int n0, Rt12 T, float4* out, int* d_n) {
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    float4 c = M.l0_cent[pos];
    float v[3] = {c.x, c.y, c.z}, r[3];
    mat3_vec(T.R, v, r);
    out[pos] = make_float4(r[0] + T.t[0], r[1] + T.t[1], r[2] + T.t[2], c.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *d_n = n0;
}

// order the purged parents by the position of their key in affected_L1 (= first touching point), lay out
// the erase sequence (children in child-set order) and drop the hash entries
__global__ void __launch_bounds__(1024) k_purge_order(MapDev M, int* us, const int* plist, const unsigned int* pfirst, int* pord, int* poff) {
  __shared__ int sm[40];
  const int P = us[US_NPURGE];
  if (P == 0) { if (threadIdx.x == 0) us[US_KPURGE] = 0; return; }
  for (int t = threadIdx.x; t < P; t += blockDim.x) {
    unsigned int f = pfirst[t]; int r = 0;
    for (int u = 0; u < P; ++u) r += (pfirst[u] < f);
    pord[r] = plist[t];
  }
  __syncthreads();
  int base = 0;
  for (int t0 = 0; t0 < P; t0 += blockDim.x) {
    int t = t0 + threadIdx.x;
    int c = t < P ? (int)M.l1_meta[pord[t]].nchild : 0, tot;
    int e = block_excl_scan(c, &tot, sm);
    if (t < P) poff[t] = base + e;
    base += tot;
  }
  if (threadIdx.x == 0) us[US_KPURGE] = base;
}
__global__ void k_purge_seq(MapDev M, int* us, const int* pord, const int* poff, int* seq_pos) {
  const int P = us[US_NPURGE];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P; t += gridDim.x * blockDim.x) {
    int s1 = pord[t];
    L1Meta* mt = &M.l1_meta[s1];
    int px, py, pz;
    morton_key(M.l1_tab[s1].key & KEY_MASK, px, py, pz);
    int N = mt->nchild, o = poff[t];
    for (int ci = 0; ci < N; ++ci) {
      int code = mt->child[ci];
      int s0 = l0_find(M, key_morton(px * M.factor + code % 3, py * M.factor + (code / 3) % 3, pz * M.factor + code / 9));
      int pos = -1;
      if (s0 >= 0) { pos = (int)M.l0_tab[s0].pos; M.l0_tab[s0].key = KEY_TOMB; }
      else atomicOr(&us[US_ERR], ERR_INTERNAL);
      seq_pos[o + ci] = pos;
    }
    M.l1_tab[s1].key = KEY_TOMB;
    mt->nchild = 0;
    atomicSub(&M.ctr[CT_N1], 1); atomicAdd(&M.ctr[CT_TOMB1], 1);
  }
}
// Replay of k swap-with-last erases (in the given order) on a dense vector of size n, on indices only.
// seq_pos[t-1]: ORIGINAL position of the t-th erased element.  Output fill[t] (at aux[3k+1+t]): original position of
// the survivor that finally sits in the hole the t-th erase leaves below the new size s = n - k.
//   loc[t]   : where the t-th erased element currently sits (>= s: tail position; -h: in the hole of erase h)
//   occ*[q-s]: current occupant of tail position q (original position id, its erase time or 0)
// Fast path (no erased element lives in the tail [s, n)): erase t moves original element n-t into hole t.
__global__ void k_swap_erase_sim(const int* us, int k_slot, int n_fixed, const int* seq_pos, int* aux, int aux_cap_k) {
  extern __shared__ int smem[];
  const int k = us[k_slot];
  if (k == 0) return;
  const int n = n_fixed >= 0 ? n_fixed : us[US_N0] + us[US_NNEW];
  const int s = n - k;
  int tail = 0;
  for (int t = threadIdx.x; t < k; t += blockDim.x) tail |= (seq_pos[t] >= s);
  if (!__syncthreads_or(tail)) {
    for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) aux[3 * k + 1 + t] = n - t;
    return;
  }
  const bool in_smem = (k <= aux_cap_k);
  int* loc = in_smem ? smem : aux;               // k+1
  int* occ_id = loc + (k + 1);                   // k
  int* occ_t = occ_id + k;                       // k
  int* fill = occ_t + k;                         // k+1
  for (int q = threadIdx.x; q < k; q += blockDim.x) { occ_id[q] = s + q; occ_t[q] = 0; }
  __syncthreads();
  for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) {
    int P = seq_pos[t - 1];
    if (P >= s) { loc[t] = P; occ_t[P - s] = t; } else loc[t] = -t;
    fill[t] = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t <= k; ++t) {
      int b = n - t;
      int l = loc[t];
      int yid = occ_id[b - s], yt = occ_t[b - s];
      if (l >= 0) { if (l != b) { occ_id[l - s] = yid; occ_t[l - s] = yt; if (yt) loc[yt] = l; } }
      else { fill[-l] = yid; if (yt) loc[yt] = l; }
    }
  }
  __syncthreads();
  if (in_smem) for (int t = 1 + threadIdx.x; t <= k; t += blockDim.x) aux[3 * k + 1 + t] = fill[t];
}
__global__ void k_purge_apply(MapDev M, int* us, const int* seq_pos, const int* aux) {
  const int k = us[US_KPURGE];
  if (k == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) us[US_N0] = us[US_N0] + us[US_NNEW]; return; }
  const int n = us[US_N0] + us[US_NNEW];
  const int s = n - k;
  const int* fill = aux + 3 * k + 1;
  for (int t = 1 + blockIdx.x * blockDim.x + threadIdx.x; t <= k; t += gridDim.x * blockDim.x) {
    int dst = seq_pos[t - 1];
    if (dst < 0 || dst >= s) continue;
    int src = fill[t];
    uint32_t sl = M.l0_slot[src];
    M.l0_cent[dst] = M.l0_cent[src];
    M.l0_key[dst] = M.l0_key[src];
    M.l0_slot[dst] = sl;
    M.l0_tab[sl].pos = (uint32_t)dst;
  }
}
__global__ void k_upd_finish(MapDev M, int* us, int purge_ran) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int n;
  if (us[US_ERR] & ERR_CAP) n = us[US_N0];
  else if (purge_ran) { int k = us[US_KPURGE]; n = (k == 0) ? us[US_N0] : us[US_N0] + us[US_NNEW] - k; if (k) atomicAdd(&M.ctr[CT_TOMB0], k); }
  else n = us[US_N0] + us[US_NNEW];
  M.ctr[CT_N0] = n;
  M.ctr[CT_ERR] = us[US_ERR];
}

// ---- table maintenance --------------------------------------------------------------------------------------
__global__ void k_l0_reinsert(MapDev M, int n0) {
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    bool ins;
    int s0 = l0_find_or_insert(M, M.l0_key[pos], &ins);
    if (s0 >= 0) { M.l0_tab[s0].pos = (uint32_t)pos; M.l0_slot[pos] = (uint32_t)s0; }
  }
}
__global__ void k_l1_reinsert(MapDev Mnew, const L1Entry* oldtab, const L1Meta* oldmeta, int oldcap) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < oldcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = oldtab[s].key;
    if (k == KEY_EMPTY || k == KEY_TOMB) continue;
    bool ins;
    int ns = l1_find_or_insert(Mnew, k & KEY_MASK, &ins);
    if (ns < 0) continue;
    L1Entry e = oldtab[s];
    Mnew.l1_tab[ns].n[0] = e.n[0]; Mnew.l1_tab[ns].n[1] = e.n[1]; Mnew.l1_tab[ns].n[2] = e.n[2];
    Mnew.l1_tab[ns].c[0] = e.c[0]; Mnew.l1_tab[ns].c[1] = e.c[1]; Mnew.l1_tab[ns].c[2] = e.c[2];
    Mnew.l1_meta[ns] = oldmeta[s];
    Mnew.l1_tab[ns].key = k;
  }
}
__global__ void k_fill_int(int* p, size_t n, int v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- exports ----------------------------------------------------------------------------------------------------
__global__ void k_export_l0(MapDev M, int n0, float* xyz, int* keys, int* counts) {
  for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n0; pos += gridDim.x * blockDim.x) {
    float4 c = M.l0_cent[pos];
    xyz[pos * 3] = c.x; xyz[pos * 3 + 1] = c.y; xyz[pos * 3 + 2] = c.z;
    if (keys) { int x, y, z; morton_key(M.l0_key[pos], x, y, z); keys[pos * 3] = x; keys[pos * 3 + 1] = y; keys[pos * 3 + 2] = z; }
    if (counts) counts[pos] = __float_as_int(c.w);
  }
}
__global__ void k_export_l1(MapDev M, int* counter, int cap, int surfels_only, int* keys, int* nchild, int* children, int* has, float* normal,
                            float* centroid, float* planarity, int* last) {
  const int tcap = 1 << M.l1_log2cap;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < tcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = M.l1_tab[s].key;
    if (k == KEY_EMPTY || k == KEY_TOMB) continue;
    if (surfels_only && !(k & SURFEL_BIT)) continue;
    int o = atomicAdd(counter, 1);
    if (o >= cap) continue;
    const L1Meta* mt = &M.l1_meta[s];
    int px, py, pz;
    morton_key(k & KEY_MASK, px, py, pz);
    if (keys) { keys[o * 3] = px; keys[o * 3 + 1] = py; keys[o * 3 + 2] = pz; }
    if (nchild) nchild[o] = mt->nchild;
    if (children) for (int ci = 0; ci < 27; ++ci) {
      int code = ci < mt->nchild ? mt->child[ci] : 0;
      children[(o * 27 + ci) * 3] = ci < mt->nchild ? px * M.factor + code % 3 : 0;
      children[(o * 27 + ci) * 3 + 1] = ci < mt->nchild ? py * M.factor + (code / 3) % 3 : 0;
      children[(o * 27 + ci) * 3 + 2] = ci < mt->nchild ? pz * M.factor + code / 9 : 0;
    }
    if (has) has[o] = (k & SURFEL_BIT) ? 1 : 0;
    if (normal) for (int a = 0; a < 3; ++a) normal[o * 3 + a] = M.l1_tab[s].n[a];
    if (centroid) for (int a = 0; a < 3; ++a) centroid[o * 3 + a] = M.l1_tab[s].c[a];
    if (planarity) planarity[o] = mt->planarity;
    if (last) last[o] = mt->last_child_count;
  }
}
__global__ void k_count_surfels(MapDev M, int* out) {
  const int tcap = 1 << M.l1_log2cap;
  int c = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < tcap; s += gridDim.x * blockDim.x) {
    unsigned long long k = M.l1_tab[s].key;
    if (k != KEY_EMPTY && k != KEY_TOMB && (k & SURFEL_BIT)) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- host side ------------------------------------------------------------------------------------------------------
static int fill_int(cudaStream_t st, int* p, size_t n, int v) {
  if (!n) return 0;
  int blocks = (int)((n + 1023) / 1024); if (blocks > 2368) blocks = 2368;
  k_fill_int<<<blocks, 1024, 0, st>>>(p, n, v);
  return 1;
}
template <class T> static int dmalloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu B) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return B2LO_E_NOMEM; }
  return B2LO_OK;
}
static int ceil_log2(size_t v) { int l = 4; while ((1ull << l) < v) ++l; return l; }

static int alloc_l0_table(b2lo_map* m, int log2cap) {
  MapDev& d = m->d;
  int rc;
  size_t cap = 1ull << log2cap;
  if ((rc = dmalloc(&d.l0_tab, cap))) return rc;
  if ((rc = dmalloc(&d.t0_first, cap))) return rc;
  if ((rc = dmalloc(&d.t0_cnt, cap))) return rc;
  if ((rc = dmalloc(&d.t0_head, cap))) return rc;
  d.l0_log2cap = log2cap; m->tcap0 = cap;
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(d.l0_tab, 0xFF, cap * sizeof(L0Entry), st));
  m->ctx->launches += fill_int(st, d.t0_first, cap, INT_MAX);
  B2_CUDA(cudaMemsetAsync(d.t0_cnt, 0, cap * sizeof(int), st));
  B2_CUDA(cudaMemsetAsync(d.t0_head, 0xFF, cap * sizeof(int), st));
  return B2LO_OK;
}
static int alloc_l1_table(b2lo_map* m, int log2cap) {
  MapDev& d = m->d;
  int rc;
  size_t cap = 1ull << log2cap;
  if ((rc = dmalloc(&d.l1_tab, cap))) return rc;
  if ((rc = dmalloc(&d.l1_meta, cap))) return rc;
  if ((rc = dmalloc(&d.t1_first, cap))) return rc;
  if ((rc = dmalloc(&d.t1_head, cap))) return rc;
  d.l1_log2cap = log2cap; m->tcap1 = cap;
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(d.l1_tab, 0xFF, cap * sizeof(L1Entry), st));
  B2_CUDA(cudaMemsetAsync(d.l1_meta, 0, cap * sizeof(L1Meta), st));
  m->ctx->launches += fill_int(st, d.t1_first, cap, INT_MAX);
  B2_CUDA(cudaMemsetAsync(d.t1_head, 0xFF, cap * sizeof(int), st));
  return B2LO_OK;
}

int map_refresh_counts(b2lo_map* m) {
  b2lo_ctx* ctx = m->ctx;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts, m->d.ctr, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  m->n0 = (size_t)ctx->h_counts[CT_N0]; m->n1 = (size_t)ctx->h_counts[CT_N1];
  m->tomb0 = (size_t)ctx->h_counts[CT_TOMB0]; m->tomb1 = (size_t)ctx->h_counts[CT_TOMB1];
  return B2LO_OK;
}

// make room for `need_l0` dense voxels and an update of `need_upd` points; rebuild tables that are too full.
// Host counts (n0, n1, tombstones) must be current.
int map_reserve(b2lo_map* m, size_t need_l0, size_t need_upd) {
  b2lo_ctx* ctx = m->ctx;
  MapDev& d = m->d;
  cudaStream_t st = ctx->stream;
  int rc;
  if (need_l0 > d.l0_cap) {  // grow the dense vectors (copy live prefix)
    size_t ncap = d.l0_cap ? d.l0_cap : 1024;
    while (ncap < need_l0) ncap *= 2;
    float4* nc = nullptr; unsigned long long* nk = nullptr; uint32_t* ns = nullptr;
    if ((rc = dmalloc(&nc, ncap)) || (rc = dmalloc(&nk, ncap)) || (rc = dmalloc(&ns, ncap))) return rc;
    if (m->n0) {
      B2_CUDA(cudaMemcpyAsync(nc, d.l0_cent, m->n0 * sizeof(float4), cudaMemcpyDeviceToDevice, st));
      B2_CUDA(cudaMemcpyAsync(nk, d.l0_key, m->n0 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
      B2_CUDA(cudaMemcpyAsync(ns, d.l0_slot, m->n0 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));
    cudaFree(d.l0_cent); cudaFree(d.l0_key); cudaFree(d.l0_slot);
    d.l0_cent = nc; d.l0_key = nk; d.l0_slot = ns; d.l0_cap = (uint32_t)ncap;
    if ((rc = dmalloc(&m->c_flag, ncap)) || (rc = dmalloc(&m->c_removed, ncap)) || (rc = dmalloc(&m->c_aux, 4 * ncap + 16)) ||
        (rc = dmalloc(&m->c_l1work, ncap)) || (rc = dmalloc(&m->c_blkcnt, ncap / 1024 + 2)) || (rc = dmalloc(&m->c_blkoff, ncap / 1024 + 2)))
      return rc;
  }
  // L0 hash: keep (live + tombstones + incoming) under half the table
  if ((m->n0 + m->tomb0 + need_upd) * 2 > m->tcap0) {
    int l2 = ceil_log2((m->n0 + need_upd) * 4);
    if (l2 < d.l0_log2cap) l2 = d.l0_log2cap;
    B2_CUDA(cudaStreamSynchronize(st));
    if ((rc = alloc_l0_table(m, l2))) return rc;
    if (m->n0) { int blocks = (int)((m->n0 + 255) / 256); if (blocks > 2368) blocks = 2368; k_l0_reinsert<<<blocks, 256, 0, st>>>(d, (int)m->n0); ctx->launches++; }
    B2_CUDA(cudaMemsetAsync(d.ctr + CT_TOMB0, 0, sizeof(int), st));
    m->tomb0 = 0;
  }
  if ((m->n1 + m->tomb1 + need_upd) * 2 > m->tcap1) {
    int l2 = ceil_log2((m->n1 + need_upd) * 4);
    if (l2 < d.l1_log2cap) l2 = d.l1_log2cap;
    B2_CUDA(cudaStreamSynchronize(st));
    L1Entry* oldtab = d.l1_tab; L1Meta* oldmeta = d.l1_meta; size_t oldcap = m->tcap1;
    d.l1_tab = nullptr; d.l1_meta = nullptr;
    if ((rc = alloc_l1_table(m, l2))) return rc;
    if (oldtab && m->n1) { int blocks = (int)((oldcap + 255) / 256); if (blocks > 2368) blocks = 2368; k_l1_reinsert<<<blocks, 256, 0, st>>>(d, oldtab, oldmeta, (int)oldcap); ctx->launches++; }
    B2_CUDA(cudaMemsetAsync(d.ctr + CT_TOMB1, 0, sizeof(int), st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (oldtab) cudaFree(oldtab);
    if (oldmeta) cudaFree(oldmeta);
    m->tomb1 = 0;
  }
  if (need_upd > m->upd_cap) {
    size_t ncap = m->upd_cap ? m->upd_cap : 4096;
    while (ncap < need_upd) ncap *= 2;
    B2_CUDA(cudaStreamSynchronize(st));
    if ((rc = dmalloc(&m->u_pts, ncap)) || (rc = dmalloc(&m->u_pslot, ncap)) || (rc = dmalloc(&m->u_next, ncap)) || (rc = dmalloc(&m->u_isnew, ncap)) ||
        (rc = dmalloc(&m->u_newrank, ncap)) || (rc = dmalloc(&m->u_l1slot, ncap)) || (rc = dmalloc(&m->u_next1, ncap)))
      return rc;
    m->a_log2cap = ceil_log2(ncap * 2);
    if ((rc = dmalloc(&m->a_tab, (size_t)1 << m->a_log2cap)) || (rc = dmalloc(&m->a_list, ncap * 4 + 16))) return rc;
    // purge scratch: up to 27 children per affected parent
    m->p_cap = ncap * 27;
    if ((rc = dmalloc(&m->p_seq, m->p_cap + 16)) || (rc = dmalloc(&m->p_aux, m->p_cap * 4 + 16))) return rc;
    m->upd_cap = ncap;
  }
  return B2LO_OK;
}

static int grid_for(size_t n, int threads) { size_t b = (n + threads - 1) / threads; if (b < 1) b = 1; if (b > 1184) b = 1184; return (int)b; }

// UpdateVoxelMap on a world-frame cloud already on the device (float4 stream).  n_cap = host-known upper
// bound of *d_n.  Ends with a counter read-back (one synchronisation).
constexpr int SIM_SMEM_K = 12000;  // (4k+2) ints <= 192 KB of dynamic shared memory
constexpr int SIM_SMEM_BYTES = (4 * SIM_SMEM_K + 2) * (int)sizeof(int);

int map_update_dev(b2lo_map* m, const float4* d_world, const int* d_n, size_t n_cap, const float sensor[3], float radius_sq, int rehash) {
  b2lo_ctx* ctx = m->ctx;
  if (n_cap == 0) return B2LO_S_EMPTY;
  int rc = map_reserve(m, m->n0 + n_cap, n_cap);
  if (rc) return rc;
  MapDev& d = m->d;
  cudaStream_t st = ctx->stream;
  int* us = m->u_state;
  if (!ctx->sim_attr_set) {
    B2_CUDA(cudaFuncSetAttribute(k_swap_erase_sim, cudaFuncAttributeMaxDynamicSharedMemorySize, SIM_SMEM_BYTES));
    ctx->sim_attr_set = true;
  }
  prof_begin(ctx, PS_MAP);
  B2_CUDA(cudaMemsetAsync(us, 0, US_COUNT * sizeof(int), st));
  const int n0 = (int)m->n0;
  if (n0 > 0 && !rehash) {
    int tiles = (n0 + 1023) / 1024;
    int g = tiles > 1184 ? 1184 : tiles;
    k_cull_mark<<<g, 1024, 0, st>>>(d, n0, sensor[0], sensor[1], sensor[2], radius_sq, m->c_flag, m->c_blkcnt);
    k_cull_scan<<<1, 1024, 0, st>>>(n0, m->c_blkcnt, m->c_blkoff, us);
    k_cull_lists<<<g, 1024, 0, st>>>(d, n0, m->c_flag, m->c_blkoff, us, m->c_removed, m->c_l1work);
    k_cull_unregister<<<grid_for(n0 / 8 + 1, 128), 128, 0, st>>>(d, m->c_flag, us, m->c_l1work);
    k_swap_erase_sim<<<1, 256, SIM_SMEM_BYTES, st>>>(us, US_K, n0, m->c_removed, m->c_aux, SIM_SMEM_K);
    k_cull_move<<<grid_for(n0 / 4 + 1, 256), 256, 0, st>>>(d, us, m->c_removed, m->c_aux);
    ctx->launches += 6;
  }
  B2_CUDA(cudaMemsetAsync(m->a_tab, 0xFF, sizeof(FEntry) << m->a_log2cap, st));
  int gm = grid_for(n_cap, 256);
  int* plist = m->a_list; unsigned int* pfirst = reinterpret_cast<unsigned int*>(m->a_list + m->upd_cap);
  int* pord = m->a_list + 2 * m->upd_cap; int* poff = m->a_list + 3 * m->upd_cap;
  k_ins_probe<<<gm, 256, 0, st>>>(d, d_world, d_n, us, m->u_pslot, m->u_next, m->a_tab, m->a_log2cap);
  k_ins_apply<<<gm, 256, 0, st>>>(d, d_world, d_n, m->u_pslot, m->u_next, m->u_isnew, m->u_pts, rehash);
  k_ins_scan<<<1, 1024, 0, st>>>(d, d_n, m->u_isnew, m->u_newrank, us);
  k_ins_place<<<gm, 256, 0, st>>>(d, d_n, us, m->u_pslot, m->u_isnew, m->u_newrank, m->u_pts, m->u_l1slot, m->u_next1);
  k_reg_children<<<gm, 256, 0, st>>>(d, d_n, us, m->u_pslot, m->u_isnew, m->u_newrank, m->u_l1slot, m->u_next1);
  ctx->launches += 5;
  int purge = 0;
  if (rehash) {
    // ApplyTransformAndRehash always ends in RecomputeAllSurfels (VoxelMap.cpp:301), whatever compute_surfels says
    k_surfel_all<<<grid_for(m->tcap1, 128), 128, 0, st>>>(d);
    ctx->launches += 1;
  } else if (d.compute_surfels) {
    purge = 1;
    k_surfel<<<grid_for((size_t)1 << m->a_log2cap, 128), 128, 0, st>>>(d, m->a_tab, m->a_log2cap, us, plist, pfirst);
    k_purge_order<<<1, 1024, 0, st>>>(d, us, plist, pfirst, pord, poff);
    k_purge_seq<<<grid_for(n_cap, 128), 128, 0, st>>>(d, us, pord, poff, m->p_seq);
    k_swap_erase_sim<<<1, 256, SIM_SMEM_BYTES, st>>>(us, US_KPURGE, -1, m->p_seq, m->p_aux, SIM_SMEM_K);
    k_purge_apply<<<grid_for(n_cap, 256), 256, 0, st>>>(d, us, m->p_seq, m->p_aux);
    ctx->launches += 5;
  }
  k_upd_finish<<<1, 32, 0, st>>>(d, us, purge);
  prof_end(ctx);
  ctx->launches += 1;
  B2_CUDA(cudaGetLastError());
  rc = map_refresh_counts(m);
  if (rc) return rc;
  int err = ctx->h_counts[CT_ERR];
  if (err & ERR_CAP) { set_error("voxel map capacity exceeded during update"); return B2LO_E_CAPACITY; }
  if (err & ERR_INTERNAL) { set_error("voxel map internal inconsistency (child without L0 voxel)"); return B2LO_E_CAPACITY; }
  if (err & ERR_RANGE) { set_error("point outside the 21-bit voxel key domain was skipped"); return B2LO_E_RANGE; }
  return B2LO_OK;
}

}  // namespace b2

// ======================================================================================================================
using namespace b2;

extern "C" int b2lo_map_create(b2lo_ctx* ctx, float voxel_size, int hierarchy_factor, float planarity_threshold, int compute_surfels,
                               size_t l0_capacity_hint, b2lo_map** out) {
  if (!ctx || !out) return B2LO_E_ARG;
  if (!(voxel_size > 0.0f)) { set_error("Voxel size must be positive"); return B2LO_E_ARG; }              // VoxelMap.cpp:28-30
  if (hierarchy_factor != 3) { set_error("Hierarchy factor must be 3 in this build (positive odd in the reference)"); return B2LO_E_ARG; }
  cudaSetDevice(ctx->device);
  b2lo_map* m = new b2lo_map();
  m->ctx = ctx;
  MapDev& d = m->d;
  d.voxel = voxel_size; d.factor = hierarchy_factor; d.scale1 = voxel_size * (float)hierarchy_factor;     // VoxelMap.cpp:51-52
  d.planarity_thr = planarity_threshold; d.compute_surfels = compute_surfels ? 1 : 0;
  if (cudaMalloc(&d.ctr, 16 * sizeof(int)) != cudaSuccess || cudaMalloc(&m->u_state, US_COUNT * sizeof(int)) != cudaSuccess) { delete m; return B2LO_E_NOMEM; }
  cudaMemsetAsync(d.ctr, 0, 16 * sizeof(int), ctx->stream);
  size_t hint = l0_capacity_hint ? l0_capacity_hint : (1u << 16);
  int rc = alloc_l0_table(m, ceil_log2(hint * 4));
  if (!rc) rc = alloc_l1_table(m, ceil_log2(hint));
  if (!rc) rc = map_reserve(m, hint, 4096);
  if (rc) { b2lo_map_destroy(m); return rc; }
  cudaStreamSynchronize(ctx->stream);
  *out = m;
  return B2LO_OK;
}

extern "C" int b2lo_map_destroy(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  MapDev& d = m->d;
  void* ptrs[] = {d.l0_cent, d.l0_key, d.l0_slot, d.l0_tab, d.l1_tab, d.l1_meta, d.t0_first, d.t0_cnt, d.t0_head, d.t1_first, d.t1_head, d.ctr,
                  m->u_pts, m->u_pslot, m->u_next, m->u_isnew, m->u_newrank, m->u_l1slot, m->u_next1, m->a_tab, m->a_list, m->c_flag, m->c_blkcnt,
                  m->c_blkoff, m->c_removed, m->c_aux, m->c_l1work, m->p_seq, m->p_aux, m->u_state};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete m;
  return B2LO_OK;
}

extern "C" int b2lo_map_clear(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  cudaSetDevice(m->ctx->device);
  cudaStream_t st = m->ctx->stream;
  B2_CUDA(cudaMemsetAsync(m->d.l0_tab, 0xFF, m->tcap0 * sizeof(L0Entry), st));
  B2_CUDA(cudaMemsetAsync(m->d.l1_tab, 0xFF, m->tcap1 * sizeof(L1Entry), st));
  B2_CUDA(cudaMemsetAsync(m->d.l1_meta, 0, m->tcap1 * sizeof(L1Meta), st));
  B2_CUDA(cudaMemsetAsync(m->d.ctr, 0, 16 * sizeof(int), st));
  B2_CUDA(cudaStreamSynchronize(st));
  m->n0 = m->n1 = m->tomb0 = m->tomb1 = 0;
  return B2LO_OK;
}
extern "C" int b2lo_map_set_planarity_threshold(b2lo_map* m, float thr) { if (!m) return B2LO_E_ARG; m->d.planarity_thr = thr; return B2LO_OK; }
extern "C" int b2lo_map_set_compute_surfels(b2lo_map* m, int on) { if (!m) return B2LO_E_ARG; m->d.compute_surfels = on ? 1 : 0; return B2LO_OK; }

extern "C" int b2lo_map_update(b2lo_map* m, const float* world_xyz, size_t n, size_t stride_floats, const double sensor[3], double max_distance) {
  if (!m || !sensor) return B2LO_E_ARG;
  if (!world_xyz || n == 0) return B2LO_S_EMPTY;  // VoxelMap.cpp:134-136
  if (stride_floats < 3) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  int rc = ctx_reserve_points(ctx, n);
  if (rc) return rc;
  rc = ctx_stage_h2d(ctx, world_xyz, n, stride_floats, 1, ctx->d_world, ctx->d_nquery);
  if (rc) return rc;
  float sf[3] = {(float)sensor[0], (float)sensor[1], (float)sensor[2]};  // sensor_position.cast<float>()  (VoxelMap.cpp:143)
  float r2 = (float)(max_distance * max_distance);                        // (:144)
  return map_update_dev(m, ctx->d_world, ctx->d_nquery, n, sf, r2);
}

extern "C" int b2lo_map_counts(b2lo_map* m, size_t* l0, size_t* l1, size_t* surfels) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  b2lo_ctx* ctx = m->ctx;
  if (l0) *l0 = m->n0;
  if (l1) *l1 = m->n1;
  if (surfels) {
    cudaSetDevice(ctx->device);
    B2_CUDA(cudaMemsetAsync(m->d.ctr + CT_SURF, 0, sizeof(int), ctx->stream));
    k_count_surfels<<<grid_for(m->tcap1, 256), 256, 0, ctx->stream>>>(m->d, m->d.ctr + CT_SURF);
    ctx->launches++;
    B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 8, m->d.ctr + CT_SURF, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    *surfels = (size_t)ctx->h_counts[8];
  }
  return B2LO_OK;
}

extern "C" int b2lo_map_export_l0(b2lo_map* m, float* xyz, int* keys, int* counts, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  b2lo_ctx* ctx = m->ctx;
  *n = m->n0;
  if (m->n0 == 0) return B2LO_OK;
  if (!xyz || cap < m->n0) { set_error("export_l0: buffer too small (%zu < %zu)", cap, m->n0); return B2LO_E_CAPACITY; }
  cudaSetDevice(ctx->device);
  size_t n0 = m->n0;
  float* dxyz = nullptr; int* dk = nullptr; int* dc = nullptr;
  B2_CUDA(cudaMalloc(&dxyz, n0 * 3 * sizeof(float)));
  if (keys) B2_CUDA(cudaMalloc(&dk, n0 * 3 * sizeof(int)));
  if (counts) B2_CUDA(cudaMalloc(&dc, n0 * sizeof(int)));
  k_export_l0<<<grid_for(n0, 256), 256, 0, ctx->stream>>>(m->d, (int)n0, dxyz, dk, dc);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(xyz, dxyz, n0 * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (keys) B2_CUDA(cudaMemcpyAsync(keys, dk, n0 * 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (counts) B2_CUDA(cudaMemcpyAsync(counts, dc, n0 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(dxyz); if (dk) cudaFree(dk); if (dc) cudaFree(dc);
  return B2LO_OK;
}

static int export_l1_common(b2lo_map* m, int surfels_only, int* keys, int* nchild, int* children, int* has, float* normal, float* centroid,
                            float* planarity, int* last, size_t cap, size_t* n) {
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  size_t n1 = m->n1;
  if (cap > n1) cap = n1;
  int *dk = nullptr, *dn = nullptr, *dch = nullptr, *dh = nullptr, *dl = nullptr, *dcount = nullptr;
  float *dnr = nullptr, *dce = nullptr, *dp = nullptr;
  size_t c1 = cap ? cap : 1;
  B2_CUDA(cudaMalloc(&dcount, sizeof(int)));
  B2_CUDA(cudaMemsetAsync(dcount, 0, sizeof(int), ctx->stream));
  if (keys) B2_CUDA(cudaMalloc(&dk, c1 * 3 * sizeof(int)));
  if (nchild) B2_CUDA(cudaMalloc(&dn, c1 * sizeof(int)));
  if (children) B2_CUDA(cudaMalloc(&dch, c1 * 81 * sizeof(int)));
  if (has) B2_CUDA(cudaMalloc(&dh, c1 * sizeof(int)));
  if (last) B2_CUDA(cudaMalloc(&dl, c1 * sizeof(int)));
  if (normal) B2_CUDA(cudaMalloc(&dnr, c1 * 3 * sizeof(float)));
  if (centroid) B2_CUDA(cudaMalloc(&dce, c1 * 3 * sizeof(float)));
  if (planarity) B2_CUDA(cudaMalloc(&dp, c1 * sizeof(float)));
  k_export_l1<<<grid_for(m->tcap1, 256), 256, 0, ctx->stream>>>(m->d, dcount, (int)cap, surfels_only, dk, dn, dch, dh, dnr, dce, dp, dl);
  ctx->launches++;
  B2_CUDA(cudaMemcpyAsync(ctx->h_counts + 9, dcount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  size_t got = (size_t)ctx->h_counts[9];
  *n = got;
  size_t cp = got < cap ? got : cap;
  if (cp) {
    if (keys) B2_CUDA(cudaMemcpy(keys, dk, cp * 3 * sizeof(int), cudaMemcpyDeviceToHost));
    if (nchild) B2_CUDA(cudaMemcpy(nchild, dn, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (children) B2_CUDA(cudaMemcpy(children, dch, cp * 81 * sizeof(int), cudaMemcpyDeviceToHost));
    if (has) B2_CUDA(cudaMemcpy(has, dh, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (last) B2_CUDA(cudaMemcpy(last, dl, cp * sizeof(int), cudaMemcpyDeviceToHost));
    if (normal) B2_CUDA(cudaMemcpy(normal, dnr, cp * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (centroid) B2_CUDA(cudaMemcpy(centroid, dce, cp * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (planarity) B2_CUDA(cudaMemcpy(planarity, dp, cp * sizeof(float), cudaMemcpyDeviceToHost));
  }
  void* ptrs[] = {dk, dn, dch, dh, dl, dcount, dnr, dce, dp};
  for (void* p : ptrs) if (p) cudaFree(p);
  return (got > cap) ? B2LO_E_CAPACITY : B2LO_OK;
}
extern "C" int b2lo_map_export_l1(b2lo_map* m, int* keys, int* nchild, int* children, int* has_surfel, float* normal, float* centroid,
                                  float* planarity, int* last_child_count, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  return export_l1_common(m, 0, keys, nchild, children, has_surfel, normal, centroid, planarity, last_child_count, cap, n);
}
extern "C" int b2lo_map_export_surfels(b2lo_map* m, float* centroid, float* normal, float* planarity, int* l1keys, size_t cap, size_t* n) {
  if (!m || !n) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  return export_l1_common(m, 1, l1keys, nullptr, nullptr, nullptr, normal, centroid, planarity, nullptr, cap, n);
}

// RebuildKdTree (VoxelMap.cpp:420-438).  The engine needs no separate index: the L0 Z-order hash is the
// uniform grid the exact 5-NN search walks (b2lo_knn.cuh), and it is always current.
int b2::map_rebuild_knn_locked(b2lo_map* m) {
  m->knn_ready = m->n0 > 0;
  return B2LO_OK;
}
extern "C" int b2lo_map_rebuild_knn(b2lo_map* m) {
  if (!m) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  return map_rebuild_knn_locked(m);
}
extern "C" int b2lo_map_has_knn(b2lo_map* m) { return (m && m->knn_ready) ? 1 : 0; }

// ApplyTransformAndRehash (VoxelMap.cpp:264-302): transform every L0 centroid, clear both levels, re-insert in dense
// order (collisions merged weighted by point_count), rebuild L1, RecomputeAllSurfels (:304-366)
extern "C" int b2lo_map_transform_rehash(b2lo_map* m, const float T16[16]) {
  if (!m || !T16) return B2LO_E_ARG;
  std::lock_guard<std::recursive_mutex> lk(m->mu);
  b2lo_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t n0 = m->n0;
  if (n0 == 0) return B2LO_S_EMPTY;
  float4* tmp = nullptr; int* d_n = nullptr;
  B2_CUDA(cudaMalloc(&tmp, n0 * sizeof(float4)));
  B2_CUDA(cudaMalloc(&d_n, sizeof(int)));
  Rt12 T;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T.R[i * 3 + j] = T16[i * 4 + j]; T.t[i] = T16[i * 4 + 3]; }
  k_xform_l0<<<grid_for(n0, 256), 256, 0, st>>>(m->d, (int)n0, T, tmp, d_n);
  ctx->launches++;
  int rc = b2lo_map_clear(m);
  if (!rc) {
    float zero[3] = {0, 0, 0};
    rc = map_update_dev(m, tmp, d_n, n0, zero, 0.0f, 1);
  }
  cudaStreamSynchronize(st);
  cudaFree(tmp); cudaFree(d_n);
  return rc;
}
