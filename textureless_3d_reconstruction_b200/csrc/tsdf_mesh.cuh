// tsdf_mesh.cuh — K10: triangle-mesh extraction from the TSDF block volume (marching cubes).
// Included by tsdf.cu inside its anonymous namespace (uses VolDev / hash_find / the K6 halo).
//
// north_star "surface/point extraction to .ply", SURVEY §8f rank 3 (no reference code; semantics
// of Open3D VoxelBlockGrid.extract_triangle_mesh, restated as R6m in oracle/t3d_oracle.c).
//
//   M1 mesh_vertex_kernel  one CTA per block: the 11^3 (tsdf, weight) halo tile of K6 in shared
//      memory -> validity of the 9^3 cubes around the block -> per own voxel the case index and
//      the 3 "edge carries a vertex" flags -> block scan -> one atomicAdd per block reserves the
//      block's vertex and triangle ranges -> vertices (R6 formulas) written in voxel order.
//      Per voxel one 32-bit word  case | flags<<8 | rank<<16  is kept (2 KiB per block) so that
//   M2 mesh_triangle_kernel  can name any edge's vertex as  vbase[block'] + rank + popc(lower flags)
//      with one gather, the owner voxel being in the block itself or one of its 7 (+x,+y,+z)
//      neighbours.
// Bytes: M1 reads 8 B x 1331 halo voxels + writes 27 B per vertex + 2 KiB meta; M2 reads the meta
// of 8 blocks (cached) and writes 12 B per triangle.
#include "mc_tables.cuh"

struct MeshOut {
  float* xyz;
  float* nrm;
  uint8_t* rgb;
  int* tri;
  long long vcap, tcap;
  unsigned long long* counts;  // [0] vertices, [1] triangles
  unsigned* meta;              // block_capacity * 512
  int* vbase;                  // per block
  int* tbase;                  // per block
};

constexpr int CV = 9;  // cube bases -1..7 per axis

__global__ void __launch_bounds__(256)
    mesh_vertex_kernel(const __grid_constant__ VolDev v, int n_blocks, float weight_thr, float voxel_size,
                       const __grid_constant__ MeshOut o) {
  __shared__ float s_t[HALO3];
  __shared__ float s_w[HALO3];
  __shared__ unsigned char s_cv[CV * CV * CV];
  __shared__ int s_nb[27];
  __shared__ int s_warp[8];
  __shared__ int s_base[2];
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    __syncthreads();
    const int bx = v.block_keys[b * 3], by = v.block_keys[b * 3 + 1], bz = v.block_keys[b * 3 + 2];
    if (tid < 27) {
      const int dx = tid % 3 - 1, dy = (tid / 3) % 3 - 1, dz = tid / 9 - 1;
      int nidx = -1;
      if (dx == 0 && dy == 0 && dz == 0) {
        nidx = b;
      } else if (key_in_range(bx + dx, by + dy, bz + dz)) {
        const long long slot = hash_find(v, pack_key(bx + dx, by + dy, bz + dz));
        if (slot >= 0) nidx = v.hvals[slot];
      }
      if (nidx >= 0 && v.fresh[nidx]) nidx = -1;
      s_nb[tid] = nidx;
    }
    __syncthreads();
    for (int i = tid; i < HALO3; i += 256) {
      const int hx = i % HALO - 1, hy = (i / HALO) % HALO - 1, hz = i / (HALO * HALO) - 1;
      const int nx = hx < 0 ? 0 : (hx >= BLK ? 2 : 1);
      const int ny = hy < 0 ? 0 : (hy >= BLK ? 2 : 1);
      const int nz = hz < 0 ? 0 : (hz >= BLK ? 2 : 1);
      const int nidx = s_nb[nx + 3 * ny + 9 * nz];
      float t = 0.f, w = -1.f;  // w = -1 marks "block missing"
      if (nidx >= 0) {
        const int lv = (hx & 7) + 8 * (hy & 7) + 64 * (hz & 7);
        const float* blk = v.blocks + (long long)nidx * BLOCK_FLOATS;
        t = blk[lv];
        w = blk[BLK3 + lv];
      }
      s_t[i] = t;
      s_w[i] = w;
    }
    __syncthreads();
#define HIDX(x, y, z) (((x) + 1) + HALO * ((y) + 1) + HALO * HALO * ((z) + 1))
    // cube validity for bases -1..7: all 8 corners observed (missing blocks carry w = -1;
    // the comparison is written so that NaN weights fail like in the oracle)
    for (int i = tid; i < CV * CV * CV; i += 256) {
      const int cx_ = i % CV - 1, cy_ = (i / CV) % CV - 1, cz_ = i / (CV * CV) - 1;
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        ok = ok && (s_w[HIDX(cx_ + (c & 1), cy_ + ((c >> 1) & 1), cz_ + (c >> 2))] >= weight_thr);
      s_cv[i] = ok ? 1 : 0;
    }
    __syncthreads();
#define CVIDX(x, y, z) (((x) + 1) + CV * ((y) + 1) + CV * CV * ((z) + 1))
    const bool present = s_nb[13] >= 0;
    unsigned cas[2], flg[2];
    int nvert = 0, ntri = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int vi = tid * 2 + h;  // consecutive voxels per thread: ranks follow voxel order
      const int xv = vi & 7, yv = (vi >> 3) & 7, zv = vi >> 6;
      cas[h] = 0;
      flg[h] = 0;
      if (present) {
        if (s_cv[CVIDX(xv, yv, zv)]) {
          unsigned c8 = 0;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (s_t[HIDX(xv + MC_CORNER[c][0], yv + MC_CORNER[c][1], zv + MC_CORNER[c][2])] < 0.f) c8 |= 1u << c;
          if (c8 != 255u) cas[h] = c8;
        }
        const bool neg_o = s_t[HIDX(xv, yv, zv)] < 0.f;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const int ex = ax == 0, ey = ax == 1, ez = ax == 2;
          const bool neg_i = s_t[HIDX(xv + ex, yv + ey, zv + ez)] < 0.f;
          if (neg_o == neg_i) continue;
          // the 4 cubes around the edge: bases shifted by {0,-1} along the two other axes
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int a = q & 1, c = q >> 1;
            const int ox = ax == 0 ? 0 : -a, oy = ax == 1 ? 0 : (ax == 0 ? -a : -c), oz = ax == 2 ? 0 : -c;
            any = any || s_cv[CVIDX(xv + ox, yv + oy, zv + oz)];
          }
          if (any) flg[h] |= 1u << ax;
        }
        nvert += __popc(flg[h]);
        ntri += MC_NUM_TRI[cas[h]];
      }
    }
    int tot_v = 0, tot_t = 0;
    const int rank_v = cta_scan256(nvert, s_warp, &tot_v);
    (void)cta_scan256(ntri, s_warp, &tot_t);
    if (tid == 0) {
      long long vb = 0, tb = 0;
      if (tot_v) vb = (long long)atomicAdd(o.counts + 0, (unsigned long long)tot_v);
      if (tot_t) tb = (long long)atomicAdd(o.counts + 1, (unsigned long long)tot_t);
      // indices are 32-bit (PLY `uint`); a range that does not fit is dropped by the capacity test
      s_base[0] = vb + tot_v <= 0x7fffffffll ? (int)vb : 0x7fffffff;
      s_base[1] = tb + tot_t <= 0x7fffffffll ? (int)tb : 0x7fffffff;
      o.vbase[b] = s_base[0];
      o.tbase[b] = s_base[1];
    }
    __syncthreads();
    const long long vbase = s_base[0];
    int r = rank_v;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int vi = tid * 2 + h;
      o.meta[(long long)b * BLK3 + vi] = cas[h] | (flg[h] << 8) | ((unsigned)r << 16);
      if (flg[h] == 0) continue;
      const int xv = vi & 7, yv = (vi >> 3) & 7, zv = vi >> 6;
      const float t_o = s_t[HIDX(xv, yv, zv)];
      for (int ax = 0; ax < 3; ++ax) {
        if (!(flg[h] & (1u << ax))) continue;
        const long long ov = vbase + r;
        ++r;
        if (o.xyz == nullptr || ov >= o.vcap) continue;
        const int ex = ax == 0, ey = ax == 1, ez = ax == 2;
        const float t_i = s_t[HIDX(xv + ex, yv + ey, zv + ez)];
        const float ratio = __fdiv_rn(__fsub_rn(0.f, t_o), __fsub_rn(t_i, t_o));
        const float gx = (float)(bx * BLK + xv), gy = (float)(by * BLK + yv), gz = (float)(bz * BLK + zv);
        o.xyz[ov * 3 + 0] = __fmul_rn(voxel_size, ex ? __fadd_rn(gx, ratio) : gx);
        o.xyz[ov * 3 + 1] = __fmul_rn(voxel_size, ey ? __fadd_rn(gy, ratio) : gy);
        o.xyz[ov * 3 + 2] = __fmul_rn(voxel_size, ez ? __fadd_rn(gz, ratio) : gz);
        const float om = __fsub_rn(1.f, ratio);
        if (o.nrm) {
          float n[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int cx_ = c == 0, cy_ = c == 1, cz_ = c == 2;
            const float go = __fsub_rn(s_t[HIDX(xv + cx_, yv + cy_, zv + cz_)],
                                       s_t[HIDX(xv - cx_, yv - cy_, zv - cz_)]);
            const float gi = __fsub_rn(s_t[HIDX(xv + ex + cx_, yv + ey + cy_, zv + ez + cz_)],
                                       s_t[HIDX(xv + ex - cx_, yv + ey - cy_, zv + ez - cz_)]);
            n[c] = __fadd_rn(__fmul_rn(om, go), __fmul_rn(ratio, gi));
          }
          const float nn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])),
                                                __fmul_rn(n[2], n[2])));
          if (nn > 0.f) { n[0] = __fdiv_rn(n[0], nn); n[1] = __fdiv_rn(n[1], nn); n[2] = __fdiv_rn(n[2], nn); }
          o.nrm[ov * 3 + 0] = n[0]; o.nrm[ov * 3 + 1] = n[1]; o.nrm[ov * 3 + 2] = n[2];
        }
        if (o.rgb) {
          const float* blk_o = v.blocks + (long long)b * BLOCK_FLOATS;
          const int nxv = xv + ex, nyv = yv + ey, nzv = zv + ez;
          const int nsel = (nxv >= BLK ? 2 : 1) + 3 * (nyv >= BLK ? 2 : 1) + 9 * (nzv >= BLK ? 2 : 1);
          const float* blk_i = v.blocks + (long long)s_nb[nsel] * BLOCK_FLOATS;
          const int li = (nxv & 7) + 8 * (nyv & 7) + 64 * (nzv & 7);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float co = blk_o[(2 + c) * BLK3 + vi], ci = blk_i[(2 + c) * BLK3 + li];
            float m = __fadd_rn(__fmul_rn(om, co), __fmul_rn(ratio, ci));
            m = fminf(fmaxf(m, 0.f), 255.f);
            o.rgb[ov * 3 + c] = (uint8_t)(int)roundf(m);
          }
        }
      }
    }
#undef CVIDX
#undef HIDX
  }
}

__global__ void __launch_bounds__(256)
    mesh_triangle_kernel(const __grid_constant__ VolDev v, int n_blocks, const __grid_constant__ MeshOut o) {
  __shared__ int s_nb[8];     // blocks at (+dx,+dy,+dz), dx,dy,dz in {0,1}
  __shared__ int s_vb[8];
  __shared__ int s_warp[8];
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    __syncthreads();
    if (v.fresh[b]) continue;  // uniform per CTA
    const int bx = v.block_keys[b * 3], by = v.block_keys[b * 3 + 1], bz = v.block_keys[b * 3 + 2];
    if (tid < 8) {
      const int dx = tid & 1, dy = (tid >> 1) & 1, dz = tid >> 2;
      int nidx = -1;
      if (tid == 0) {
        nidx = b;
      } else if (key_in_range(bx + dx, by + dy, bz + dz)) {
        const long long slot = hash_find(v, pack_key(bx + dx, by + dy, bz + dz));
        if (slot >= 0) nidx = v.hvals[slot];
      }
      if (nidx >= 0 && v.fresh[nidx]) nidx = -1;
      s_nb[tid] = nidx;
      s_vb[tid] = nidx >= 0 ? o.vbase[nidx] : 0;
    }
    __syncthreads();
    unsigned cas[2];
    int ntri = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      cas[h] = o.meta[(long long)b * BLK3 + tid * 2 + h] & 255u;
      ntri += MC_NUM_TRI[cas[h]];
    }
    int tot = 0;
    int r = cta_scan256(ntri, s_warp, &tot);
    if (tot == 0 || o.tri == nullptr) continue;
    const long long tbase = o.tbase[b];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (cas[h] == 0) continue;
      const int vi = tid * 2 + h;
      const int xv = vi & 7, yv = (vi >> 3) & 7, zv = vi >> 6;
      const int nt = MC_NUM_TRI[cas[h]];
      for (int k = 0; k < nt; ++k) {
        const long long ot = tbase + r;
        ++r;
        if (ot >= o.tcap) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int e = MC_TRI[cas[h]][k * 3 + j];
          const int ox = xv + MC_EDGE_SHIFT[e][0], oy = yv + MC_EDGE_SHIFT[e][1], oz = zv + MC_EDGE_SHIFT[e][2];
          const int ax = MC_EDGE_SHIFT[e][3];
          const int sel = (ox >> 3) + 2 * (oy >> 3) + 4 * (oz >> 3);
          const unsigned m = o.meta[(long long)s_nb[sel] * BLK3 + (ox & 7) + 8 * (oy & 7) + 64 * (oz & 7)];
          o.tri[ot * 3 + j] = s_vb[sel] + (int)(m >> 16) + __popc((m >> 8) & ((1u << ax) - 1u));
        }
      }
    }
  }
}
