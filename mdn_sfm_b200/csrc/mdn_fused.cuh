// mdn_fused.cuh -- the fused tile kernel.  Included by mdn_loss.cu INSIDE namespace mdn (after KParams).
//
// One CTA = one 32x16 tile of one sample at one scale; both (target, source) pairs are processed by the same CTA
// so that d(loss)/d(mobile) is written exactly once.  Per pair:
//   P1  flow -> sampling coordinates -> bilinear gather of the source image over the tile + 2-pixel halo, written to
//       shared memory ALREADY reflection-padded (slot -1 holds pixel 1, slot h holds pixel h-2), so every 3x3 window
//       below is a plain box; the two pixels a thread owns keep d(warped)/d(ix,iy) and the validity bit in registers
//   P2  SSIM over the tile + 1-pixel halo: each thread slides down 3 windows of one column, sharing the horizontal
//       3-tap sums of 5 rows (separable box filter); writes the three adjoint coefficients (A,B,C) per channel
//   P3  the thread's two vertically adjacent pixels: 3x3 adjoint gather (sharing 4 rows of horizontal sums), L1
//       adjoint, chain rule to the flow; epipolar distance, post-processing, masked sums, their adjoints
// then  P4 smoothness + consistency + routing of d/dmask through the min, and the block reduction of 40 partial sums.
//
// The kernel is ISSUE-bound, not HBM-bound: everything here is about instruction count.

constexpr int RING = R2N - TN;   // halo slots of the halo-2 region

struct Smem {
  float* T;      // [3][R2N] target image, reflection padded
  float* W;      // [3][R2N] warped source image (current pair), reflection padded
  float* M;      // [3][R1N] raw mobile maps 0 / 1 and, in MIN mode, their minimum
  float* ABC;    // [9][R1N] SSIM adjoint coefficients per window: (A,B,C) x 3 channels (current pair)
  float* red;    // [nwarps][NSLOT]
};

__host__ __device__ constexpr size_t fused_smem_floats(bool photo, int nwarps) {
  return 3 * R2N + 3 * R1N + (size_t)nwarps * NSLOT + (photo ? 3 * R2N + 9 * R1N : 0);
}

template <int NV>
MDN_DEV void flush_acc(float* v, float* red, int slot_base) {
  warp_reduce_transpose<NV>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < NV) red[warp * NSLOT + slot_base + lane] = v[0];
}

// image coordinate whose value a shared-memory slot at coordinate t holds (ReflectionPad2d(1)); -1 = none
MDN_DEV int stage_index(int t, int n) {
  if (t < 0) return (t == -1) ? 1 : -1;
  if (t >= n) return (t == n) ? n - 2 : -1;
  return t;
}

// (ry, rx) of the j-th halo slot of the halo-2 region: two top rows, two bottom rows, then the side columns
MDN_DEV void ring_slot(int j, int& ry, int& rx) {
  if (j < 2 * R2W) { ry = j / R2W; rx = j - ry * R2W; }
  else if (j < 4 * R2W) { j -= 2 * R2W; ry = j / R2W; rx = j - ry * R2W; ry += TH + 2; }
  else { j -= 4 * R2W; ry = 2 + (j >> 2); int k = j & 3; rx = (k < 2) ? k : TW + k; }
}

struct PixState {          // what a thread keeps in registers about one of its two pixels, per pair
  float ddx[3], ddy[3];    // d(warped_c)/d(ix), d(warped_c)/d(iy)
  bool valid;
};

// 3x3 adjoint gather of one coefficient plane for the thread's two vertically adjacent pixels.  Q points at the
// halo-1 slot (ly0, lx): rows ly0 .. ly0+3, columns lx .. lx+2.  BORDER applies the reflection multiplicities.
template <bool BORDER>
MDN_DEV void adjoint_box(const float* Q, const float* wxm, const float (*wym)[3], float& s0, float& s1) {
  float H[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const float q0 = Q[rr * R1W], q1 = Q[rr * R1W + 1], q2 = Q[rr * R1W + 2];
    H[rr] = BORDER ? (wxm[0] * q0 + wxm[1] * q1 + wxm[2] * q2) : (q0 + q1 + q2);
  }
  if (BORDER) {
    s0 = wym[0][0] * H[0] + wym[0][1] * H[1] + wym[0][2] * H[2];
    s1 = wym[1][0] * H[1] + wym[1][1] * H[2] + wym[1][2] * H[3];
  } else {
    const float mid = H[1] + H[2];
    s0 = H[0] + mid;
    s1 = mid + H[3];
  }
}

template <bool PHOTO, bool MAPS>
__global__ void __launch_bounds__(NTHREADS, 4) fused_tile_kernel(const __grid_constant__ KParams P) {
  MDN_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, nwarps = nthr >> 5;
  const bool use_ssim = PHOTO && (P.flags & MDN_OPT_SSIM);
  const bool epi_on = (P.flags & MDN_TERM_EPIPOLAR) != 0;
  const bool smooth_on = (P.flags & MDN_TERM_SMOOTH) != 0;
  const bool consis_on = (P.flags & MDN_TERM_CONSIS) != 0;
  const bool grads = (P.flags & MDN_OPT_GRADS) != 0;
  const bool own = P.mask_mode == MDN_MASK_OWN;
  const bool shared_mask = P.mask_mode == MDN_MASK_SHARED;
  const bool need_tgt = PHOTO || smooth_on;
  const bool need_mask = epi_on || smooth_on || consis_on;

  Smem sm;
  {
    float* p = smem_raw;
    sm.T = p; p += 3 * R2N;
    sm.M = p; p += 3 * R1N;
    sm.red = p; p += nwarps * NSLOT;
    sm.W = p; sm.ABC = p;
    if (PHOTO) { sm.W = p; p += 3 * R2N; sm.ABC = p; }
  }

  // ---- which tile
  int s = 0;
#pragma unroll
  for (int k = 1; k < MDN_MAX_SCALES; ++k)
    if (k < P.n_scales && (int)blockIdx.x >= P.sc[k].tile_begin) s = k;
  const KScale& S = P.sc[s];
  int r = blockIdx.x - S.tile_begin;
  const int tiles_per_img = S.tiles_x * S.tiles_y;
  const int b = r / tiles_per_img;
  r -= b * tiles_per_img;
  const int ty = r / S.tiles_x, tx = r - ty * S.tiles_x;
  const int x0 = tx * TW, y0 = ty * TH;
  const int h = S.h, w = S.w, hw = h * w;
  // tile holds a pixel whose 3x3 adjoint gather sees a reflected tap (rows 1, h-2 / columns 1, w-2)
  const bool border = (y0 == 0) | (h - 2 >= y0 && h - 2 < y0 + TH) | (x0 == 0) | (w - 2 >= x0 && w - 2 < x0 + TW);

  // the two pixels this thread owns: same column, vertically adjacent
  const int lx = tid & 31, ly0 = (tid >> 5) * 2;
  const int px = x0 + lx;
  const bool col_in = px < w;

  for (int i = tid; i < nwarps * NSLOT; i += nthr) sm.red[i] = 0.f;

  // ---- P0: stage the target image (halo 2, reflection padded) and the raw mobile maps (halo 1) with cp.async:
  // no registers, no waiting -- the copies land while P1 computes coordinates and gathers the source image.
  if (need_tgt) {
    const float* tg = S.tgt + (size_t)b * 3 * hw;
#pragma unroll
    for (int j = 0; j < (R2N + NTHREADS - 1) / NTHREADS; ++j) {
      const int i = tid + j * NTHREADS;
      if (i < R2N) {
        const int ry = i / R2W, rx = i - ry * R2W;
        const int yy = stage_index(y0 - 2 + ry, h), xx = stage_index(x0 - 2 + rx, w);
        const bool ok = (yy | xx) >= 0;
        const float* src = tg + (ok ? yy * w + xx : 0);
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async_f32(sm.T + c * R2N + i, src + c * hw, ok);
      }
    }
  }
  if (need_mask) {
    const float* m0 = S.mob[0] + (size_t)b * hw;
    const float* m1 = shared_mask ? m0 : S.mob[1] + (size_t)b * hw;
#pragma unroll
    for (int j = 0; j < (R1N + NTHREADS - 1) / NTHREADS; ++j) {
      const int i = tid + j * NTHREADS;
      if (i < R1N) {
        const int ry = i / R1W, rx = i - ry * R1W;
        const int y = y0 - 1 + ry, x = x0 - 1 + rx;
        const bool in = (y >= 0) & (y < h) & (x >= 0) & (x < w);
        const int o = in ? y * w + x : 0;
        cp_async_f32(sm.M + i, m0 + o, in);
        if (!shared_mask) cp_async_f32(sm.M + R1N + i, m1 + o, in);
      }
    }
  }
  // which plane the 3x3 / stencil reads of a pair use: its own map (OWN), the given map (SHARED) or the minimum (MIN)
  const float* Mmin = (own | shared_mask) ? sm.M : sm.M + 2 * R1N;
  bool staged = false;   // cp.async copies completed, minimum plane built, block synchronised

  float mbar[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [pixel][mask slot] accumulated d(loss)/d(mask)

  // ---- per (target, source) pair
  for (int pair = 0; pair < P.n_pairs; ++pair) {
    float acc[PAIR_SLOTS];
#pragma unroll
    for (int k = 0; k < PAIR_SLOTS; ++k) acc[k] = 0.f;
    const float* flx = S.flow[pair] + (size_t)b * 2 * hw;
    const float* fly = flx + hw;
    const int mslot = own ? pair : 0;
    const float* Mp = own ? sm.M + pair * R1N : Mmin;
    PixState ps[2];
    float pfx[2] = {0.f, 0.f}, pfy[2] = {0.f, 0.f};   // pixel flow of the two own pixels (P1 -> P3)

    if (PHOTO) {
      const float* rf = S.ref[pair] + (size_t)b * 3 * hw;
      // -- P1: flow -> coordinates -> bilinear gather for the thread's three halo-2 slots: its own two pixels (real
      // pixels when inside the image, reflected copies otherwise) and one slot of the halo ring.  All six flow loads
      // are issued before the first use, then the 12 gathers of each slot are in flight together.
      const float* rf1 = rf + hw;
      const float* rf2 = rf1 + hw;
      int so[3], si2[3], sxx[3], syy[3];
      bool sok[3];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        syy[k] = stage_index(y0 + ly0 + k, h); sxx[k] = stage_index(px, w);
        si2[k] = (ly0 + k + 2) * R2W + lx + 2;
      }
      {
        int ry, rx;
        ring_slot(tid < RING ? tid : 0, ry, rx);
        syy[2] = stage_index(y0 - 2 + ry, h); sxx[2] = stage_index(x0 - 2 + rx, w);
        si2[2] = ry * R2W + rx;
      }
      float ffx[3], ffy[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        sok[k] = (syy[k] | sxx[k]) >= 0;
        so[k] = sok[k] ? syy[k] * w + sxx[k] : 0;
        ffx[k] = __ldg(flx + so[k]); ffy[k] = __ldg(fly + so[k]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (k == 2 && tid >= RING) break;
        const bool real = (k < 2) && (y0 + ly0 + k < h) && col_in;
        const float fx = __fmul_rn(S.sx, ffx[k]), fy = __fmul_rn(S.sy, ffy[k]);
        if (k < 2) { pfx[k] = fx; pfy[k] = fy; }
        WarpCoord wc = warp_coord((float)sxx[k], (float)syy[k], fx, fy, S.geom);
        Gather4 gt = gather_setup(wc.ix, wc.iy, h, w);
        if (k < 2) ps[k].valid = wc.valid & real;
        const float okf = sok[k] ? 1.f : 0.f;
        float v4[3][4];
        gather_fetch(rf, gt, v4[0][0], v4[0][1], v4[0][2], v4[0][3]);
        gather_fetch(rf1, gt, v4[1][0], v4[1][1], v4[1][2], v4[1][3]);
        gather_fetch(rf2, gt, v4[2][0], v4[2][1], v4[2][2], v4[2][3]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float wv = gather_value(gt, v4[c][0], v4[c][1], v4[c][2], v4[c][3]) * okf;
          if (k < 2) gather_deriv(gt, v4[c][0], v4[c][1], v4[c][2], v4[c][3], ps[k].ddx[c], ps[k].ddy[c]);
          sm.W[c * R2N + si2[k]] = wv;
          if (MAPS && real && S.warped[pair]) S.warped[pair][((size_t)b * 3 + c) * hw + so[k]] = wv;
        }
        if (MAPS && real && S.valid[pair]) S.valid[pair][(size_t)b * hw + so[k]] = wc.valid ? 1 : 0;
      }
      if (!staged) {   // first pair only: the staging copies must have landed before anybody reads T / M
        cp_async_wait_all();
        if (need_mask & !own & !shared_mask) {
#pragma unroll
          for (int j = 0; j < (R1N + NTHREADS - 1) / NTHREADS; ++j) {
            const int i = tid + j * NTHREADS;   // the slots this thread staged itself
            if (i < R1N) { const float a0 = sm.M[i], a1 = sm.M[R1N + i]; sm.M[2 * R1N + i] = (a0 <= a1) ? a0 : a1; }
          }
        }
        staged = true;
      }
      __syncthreads();

      // -- P2: SSIM; thread (column cw of the halo-1 region, group g) slides over windows rows 3g .. 3g+2
      if (use_ssim) {
        if (tid < R1W * (R1H / 3)) {
          const int g = tid / R1W, cw = tid - g * R1W;
          const int wx = x0 - 1 + cw;
          const bool colok = (wx >= 0) & (wx < w);
          const bool col_interior = (cw >= 1) & (cw <= TW);
          const float k9 = S.c_ssim * (1.f / 9.f);
#pragma unroll 1   // keep the body once in the instruction cache: the kernel is fetch-sensitive (4 CTAs in 4 different phases)
          for (int c = 0; c < 3; ++c) {
            const float* Tc = sm.T + c * R2N + (3 * g) * R2W + cw;
            const float* Wc = sm.W + c * R2N + (3 * g) * R2W + cw;
            float hx[5], hy[5], hxx[5], hyy[5], hxy[5];
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) {
              float a0 = Tc[rr * R2W], a1 = Tc[rr * R2W + 1], a2 = Tc[rr * R2W + 2];
              float b0 = Wc[rr * R2W], b1 = Wc[rr * R2W + 1], b2 = Wc[rr * R2W + 2];
              hx[rr] = a0 + a1 + a2;
              hy[rr] = b0 + b1 + b2;
              hxx[rr] = a0 * a0 + a1 * a1 + a2 * a2;
              hyy[rr] = b0 * b0 + b1 * b1 + b2 * b2;
              hxy[rr] = a0 * b0 + a1 * b1 + a2 * b2;
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const int rw = 3 * g + q;                 // window row in the halo-1 region
              const int wy = y0 - 1 + rw;
              const int i1 = rw * R1W + cw;
              // branch free: windows outside the image are evaluated on zeros and masked out
              const bool inimg = colok & (wy >= 0) & (wy < h);
              const bool inter = inimg & col_interior & (rw >= 1) & (rw <= TH);
              SsimOut so = ssim_window(hx[q] + hx[q + 1] + hx[q + 2], hy[q] + hy[q + 1] + hy[q + 2],
                                       hxx[q] + hxx[q + 1] + hxx[q + 2], hyy[q] + hyy[q + 1] + hyy[q + 2],
                                       hxy[q] + hxy[q + 1] + hxy[q + 2], grads);
              acc[SL_SSIM] += inter ? so.J : 0.f;
              if (MAPS && inter && S.ssim_map[pair]) S.ssim_map[pair][((size_t)b * 3 + c) * hw + wy * w + wx] = so.J;
              const float kk = inimg ? k9 : 0.f;
              const float A = kk * so.dmu_y, Bc = kk * 2.f * so.dY2, Cc = kk * so.dXY;
              sm.ABC[(3 * c + 0) * R1N + i1] = A;
              sm.ABC[(3 * c + 1) * R1N + i1] = Bc;
              sm.ABC[(3 * c + 2) * R1N + i1] = Cc;
            }
          }
        }
        __syncthreads();
      }
    }

    if (!PHOTO && !staged) {
      cp_async_wait_all();
      if (need_mask & !own & !shared_mask) {
#pragma unroll
        for (int j = 0; j < (R1N + NTHREADS - 1) / NTHREADS; ++j) {
          const int i = tid + j * NTHREADS;
          if (i < R1N) { const float a0 = sm.M[i], a1 = sm.M[R1N + i]; sm.M[2 * R1N + i] = (a0 <= a1) ? a0 : a1; }
        }
      }
      staged = true;
      __syncthreads();
    }
    // -- P3: the thread's two pixels: photometric adjoint -> d/dflow, epipolar forward + adjoint
    {
      float gfx[2] = {0.f, 0.f}, gfy[2] = {0.f, 0.f};
      if (PHOTO) {   // L1 term: |tgt - warped| * valid  (loss_functions.py:109-110)
        const int i2 = (ly0 + 2) * R2W + lx + 2;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float df = fabsf(sm.T[c * R2N + i2 + k * R2W] - sm.W[c * R2N + i2 + k * R2W]);
            df = ps[k].valid ? df : 0.f;     // valid already implies "real pixel"
            acc[SL_L1] += df;
            if (MAPS && S.diff[pair] && (y0 + ly0 + k < h) && col_in)
              S.diff[pair][((size_t)b * 3 + c) * hw + (y0 + ly0 + k) * w + px] = df;
          }
        }
      }
      if (PHOTO && grads) {
        float wxm[3] = {1.f, 1.f, 1.f}, wym[2][3] = {{1.f, 1.f, 1.f}, {1.f, 1.f, 1.f}};
        if (border) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            int pxx = px + d - 1;
            wxm[d] = (pxx >= 0 && pxx < w) ? (float)refl_mult(pxx, px, w) : 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              int yq = y0 + ly0 + k, pyy = yq + d - 1;
              wym[k][d] = (pyy >= 0 && pyy < h) ? (float)refl_mult(pyy, yq, h) : 0.f;
            }
          }
        }
        float gix[2] = {0.f, 0.f}, giy[2] = {0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float wb[2] = {0.f, 0.f};
          const int i2 = (ly0 + 2) * R2W + lx + 2;
          const float tv[2] = {sm.T[c * R2N + i2], sm.T[c * R2N + i2 + R2W]};
          const float wv[2] = {sm.W[c * R2N + i2], sm.W[c * R2N + i2 + R2W]};
          if (use_ssim) {
            float S3[2][3];
            const float* Q = sm.ABC + (3 * c) * R1N + ly0 * R1W + lx;   // halo-1 rows ly0 .. ly0+3, cols lx .. lx+2
            if (border) {
#pragma unroll
              for (int a = 0; a < 3; ++a) adjoint_box<true>(Q + a * R1N, wxm, wym, S3[0][a], S3[1][a]);
            } else {
#pragma unroll
              for (int a = 0; a < 3; ++a) adjoint_box<false>(Q + a * R1N, wxm, wym, S3[0][a], S3[1][a]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) wb[k] = S3[k][0] + wv[k] * S3[k][1] + tv[k] * S3[k][2];
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (ps[k].valid) wb[k] -= S.c_l1 * signf_(tv[k] - wv[k]);
            gix[k] += wb[k] * ps[k].ddx[c];
            giy[k] += wb[k] * ps[k].ddy[c];
          }
        }
        // (w-1)/2 [grid_sample] * 2 [2g-1] / (w-1) [/= w-1] * sx [scale factor]
#pragma unroll
        for (int k = 0; k < 2; ++k) { gfx[k] = gix[k] * S.sx; gfy[k] = giy[k] * S.sy; }
      }
      float Fm[9];
      float snmax = 1.f;
      if (epi_on) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Fm[k] = __ldg(S.fmat[pair] + b * 9 + k);
        if (P.post == MDN_POST_SN) {
          unsigned long long key = P.snkey[(s * P.n_pairs + pair) * P.batch + b];
          snmax = __uint_as_float((unsigned)(key >> 32));
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ly = ly0 + k, y = y0 + ly;
        if (!((y < h) & col_in)) continue;
        const int o = y * w + px;
        if (epi_on) {
          const float m = Mp[(ly + 1) * R1W + lx + 1];
          const float xf = (float)px, yf = (float)y;
          float u, v;
          if (PHOTO) { u = __fadd_rn(xf, pfx[k]); v = __fadd_rn(yf, pfy[k]); }
          else { u = __fadd_rn(xf, __fmul_rn(S.sx, __ldg(flx + o))); v = __fadd_rn(yf, __fmul_rn(S.sy, __ldg(fly + o))); }
          Epi e = epipolar_distance(Fm, xf, yf, u, v);
          float ae = fabsf(e.d);
          float dpost;
          float post = post_process(P, S, ae, snmax, o, dpost);
          float kmask = 1.f;
          if ((P.flags & (MDN_OPT_INST_MASK | MDN_OPT_CROSS_ENT)) != 0) kmask = (float)__ldg(S.inst + (size_t)b * hw + o);
          if (P.flags & MDN_OPT_INST_MASK) { post *= kmask; dpost *= kmask; }
          float bg = 1.f - m;
          float lg = __logf(bg + 1e-5f);
          float ml = m * lg;
          acc[SL_EPI] += bg * post;
          acc[SL_NT] += fabsf(ml);
          float mb = 0.f;
          if (P.flags & MDN_OPT_CROSS_ENT) {
            float l1 = __logf(m + 1e-10f), l0 = __logf(bg + 1e-10f);
            acc[SL_CE] += -(kmask * l1 + (1.f - kmask) * l0);
            mb += S.c_ce * (__fdividef(1.f - kmask, bg + 1e-10f) - __fdividef(kmask, m + 1e-10f));
          }
          if (MAPS && S.post_map[pair]) S.post_map[pair][(size_t)b * hw + o] = post;
          if (MAPS && S.ori_map[pair]) S.ori_map[pair][(size_t)b * hw + o] = (P.post == MDN_POST_SN) ? __fdiv_rn(ae, snmax) : ae;
          if (grads) {
            mb += -S.c_epi * post + S.c_nt * signf_(ml) * (lg - __fdividef(m, bg + 1e-5f));
            if (mslot) mbar[k][1] += mb; else mbar[k][0] += mb;
            float ebar = S.c_epi * bg * dpost;
            float dbar = signf_(e.d) * ebar;
            float g2 = __fdividef(dbar, e.den);
            float das = __fdividef(e.d, e.s);
            gfx[k] += g2 * e.a * S.sx;
            gfy[k] += g2 * e.b * S.sy;
            float g0 = g2 * (u - das * e.a), g1 = g2 * (v - das * e.b);
            acc[SL_GF + 0] += g0 * xf; acc[SL_GF + 1] += g0 * yf; acc[SL_GF + 2] += g0;
            acc[SL_GF + 3] += g1 * xf; acc[SL_GF + 4] += g1 * yf; acc[SL_GF + 5] += g1;
            acc[SL_GF + 6] += g2 * xf; acc[SL_GF + 7] += g2 * yf; acc[SL_GF + 8] += g2;
          }
        }
        if (grads && S.g_flow[pair]) {
          S.g_flow[pair][(size_t)b * 2 * hw + o] = gfx[k];
          S.g_flow[pair][(size_t)b * 2 * hw + hw + o] = gfy[k];
        }
      }
    }
    flush_acc<PAIR_SLOTS>(acc, sm.red, pair * PAIR_SLOTS);
    __syncthreads();   // W / ABC are reused by the next pair
  }

  // ---- P4: smoothness + consistency, then route d/dmask to the mobile maps
  if (need_mask) {
    float acc[TAIL_SLOTS];
#pragma unroll
    for (int k = 0; k < TAIL_SLOTS; ++k) acc[k] = 0.f;
    const int n_masks = own ? P.n_pairs : 1;
    // in MIN / SHARED mode the reference evaluates smooth_loss once per source frame with the SAME mask
    const float rep = own ? 1.f : (float)P.n_pairs;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int ly = ly0 + k, y = y0 + ly;
      if (!((y < h) & col_in)) continue;
      const int o = y * w + px;
      const int i1 = (ly + 1) * R1W + lx + 1, i2 = (ly + 2) * R2W + lx + 2;
      if (smooth_on) {
        // exp(-mean_c |I(x) - I(x+1)|) for the pixel pairs (x-1,x), (x,x+1), (y-1,y), (y,y+1)
        float gr = 0.f, gl = 0.f, gd = 0.f, gu = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float t0 = sm.T[c * R2N + i2];
          gr += fabsf(t0 - sm.T[c * R2N + i2 + 1]);
          gl += fabsf(sm.T[c * R2N + i2 - 1] - t0);
          gd += fabsf(t0 - sm.T[c * R2N + i2 + R2W]);
          gu += fabsf(sm.T[c * R2N + i2 - R2W] - t0);
        }
        const float third = 1.f / 3.f;
        const float ex_r = (px + 1 < w) ? __expf(-gr * third) : 0.f;
        const float ex_l = (px > 0) ? __expf(-gl * third) : 0.f;
        const float ey_d = (y + 1 < h) ? __expf(-gd * third) : 0.f;
        const float ey_u = (y > 0) ? __expf(-gu * third) : 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q < n_masks) {
            const float* Mk = own ? sm.M + q * R1N : Mmin;
            float mc = Mk[i1];
            float dr = mc - Mk[i1 + 1], dl = Mk[i1 - 1] - mc, dd = mc - Mk[i1 + R1W], du = Mk[i1 - R1W] - mc;
            acc[SL_SMX + 2 * q] += fabsf(dr) * ex_r;   // ex_r == 0 at the last column
            acc[SL_SMY + 2 * q] += fabsf(dd) * ey_d;
            mbar[k][q] += rep * (S.c_smx * (signf_(dr) * ex_r - signf_(dl) * ex_l) + S.c_smy * (signf_(dd) * ey_d - signf_(du) * ey_u));
          }
        }
      }
      const float a0 = sm.M[i1], a1 = shared_mask ? a0 : sm.M[R1N + i1];   // raw maps at this pixel
      float g0, g1;
      if (own) { g0 = mbar[k][0]; g1 = mbar[k][1]; }
      else if (shared_mask) { g0 = mbar[k][0]; g1 = 0.f; }
      else { bool first = a0 <= a1; g0 = first ? mbar[k][0] : 0.f; g1 = first ? 0.f : mbar[k][0]; }
      if (consis_on) {
        float p = __fdividef(1.f, 1.f + __expf(-20.f * (a0 - 0.5f))), q = __fdividef(1.f, 1.f + __expf(-20.f * (a1 - 0.5f)));
        float df = p - q;
        acc[SL_CONSIS] += df * df;
        g0 += S.c_consis * 40.f * df * p * (1.f - p);
        g1 -= S.c_consis * 40.f * df * q * (1.f - q);
      }
      if (grads) {
        if (S.g_mob[0]) S.g_mob[0][(size_t)b * hw + o] = g0;
        if (S.g_mob[1] && !shared_mask) S.g_mob[1][(size_t)b * hw + o] = g1;
      }
    }
    flush_acc<TAIL_SLOTS>(acc, sm.red, TAIL_BASE);
  }
  __syncthreads();
  for (int k = tid; k < NSLOT; k += nthr) {
    float t = 0.f;
    for (int wq = 0; wq < nwarps; ++wq) t += sm.red[wq * NSLOT + k];
    P.partials[(size_t)blockIdx.x * NSLOT + k] = t;
  }
}
