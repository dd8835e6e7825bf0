// scan_core.cuh -- the adding method as warp scans.
//
// The reference marches the adding method layer by layer (radiation_adding_ica_sw.F90:85-146, radiation_adding_ica_lw.F90:
// 137-263): an upward recurrence for the albedo / source of everything below a half-level, then a downward recurrence for the
// fluxes.  Both are compositions of maps with a closed associative form:
//   albedo        A(l) = R + T^2 A(l+1) / (1 - A(l+1) R)            a Moebius map of A(l+1): 2x2 matrix [T^2-R^2, R; -R, 1]
//   source/direct S(l) = alpha + beta S(l+1)                        affine, once A(l+1) is known
//   fluxes        (dir, dn)(l+1) = [t, 0; b, a] (dir, dn)(l)        lower-triangular 2x2, once A, S are known
// so a warp can own ONE (column, g-point), its 32 lanes holding LPL consecutive layers each: every lane composes the maps of
// its own layers, a warp scan (shuffles) composes across lanes, and each lane then replays its LPL layers with the reference's
// own formulae starting from the exact carry-in.  The per-layer two-stream solutions stay in registers between the upward and
// the downward pass: no adding-method state goes to memory at all.
#pragma once
#include <cuda_runtime.h>

namespace ecb {

#define ECB_FULL 0xffffffffu

struct Mob { double a, b, c, d; };   // x -> (a x + b) / (c x + d)
struct Aff { double al, be; };       // x -> al + be x
struct Tri { double t, b, a; };      // (dir, dn) -> (t dir, b dir + a dn)

__device__ __forceinline__ Mob mob_id() { Mob m; m.a = 1.0; m.b = 0.0; m.c = 0.0; m.d = 1.0; return m; }
__device__ __forceinline__ Aff aff_id() { Aff f; f.al = 0.0; f.be = 1.0; return f; }
__device__ __forceinline__ Tri tri_id() { Tri m; m.t = 1.0; m.b = 0.0; m.a = 1.0; return m; }
// x . y : apply y first, then x
__device__ __forceinline__ Mob mob_mul(const Mob& x, const Mob& y) {
  Mob r;
  r.a = x.a * y.a + x.b * y.c; r.b = x.a * y.b + x.b * y.d;
  r.c = x.c * y.a + x.d * y.c; r.d = x.c * y.b + x.d * y.d;
  return r;
}
// the adding step of one layer (reflectance R, transmittance T) appended BELOW what x already holds: x . [T^2-R^2, R; -R, 1]
__device__ __forceinline__ Mob mob_push_below(const Mob& x, double R, double T) {
  const double p = T * T - R * R;
  Mob r;
  r.a = x.a * p - x.b * R; r.b = x.a * R + x.b;
  r.c = x.c * p - x.d * R; r.d = x.c * R + x.d;
  return r;
}
__device__ __forceinline__ double mob_apply(const Mob& m, double x) { return (m.a * x + m.b) / (m.c * x + m.d); }
__device__ __forceinline__ Aff aff_mul(const Aff& x, const Aff& y) { Aff r; r.al = x.al + x.be * y.al; r.be = x.be * y.be; return r; }
__device__ __forceinline__ Tri tri_mul(const Tri& x, const Tri& y) { Tri r; r.t = x.t * y.t; r.b = x.b * y.t + x.a * y.b; r.a = x.a * y.a; return r; }

__device__ __forceinline__ Mob shfl_down(const Mob& m, int d) {
  Mob r; r.a = __shfl_down_sync(ECB_FULL, m.a, d); r.b = __shfl_down_sync(ECB_FULL, m.b, d);
  r.c = __shfl_down_sync(ECB_FULL, m.c, d); r.d = __shfl_down_sync(ECB_FULL, m.d, d); return r;
}
__device__ __forceinline__ Aff shfl_down(const Aff& m, int d) { Aff r; r.al = __shfl_down_sync(ECB_FULL, m.al, d); r.be = __shfl_down_sync(ECB_FULL, m.be, d); return r; }
__device__ __forceinline__ Aff shfl_up(const Aff& m, int d) { Aff r; r.al = __shfl_up_sync(ECB_FULL, m.al, d); r.be = __shfl_up_sync(ECB_FULL, m.be, d); return r; }
__device__ __forceinline__ Tri shfl_up(const Tri& m, int d) {
  Tri r; r.t = __shfl_up_sync(ECB_FULL, m.t, d); r.b = __shfl_up_sync(ECB_FULL, m.b, d); r.a = __shfl_up_sync(ECB_FULL, m.a, d); return r;
}

// Upward scans (layers are numbered top-down, lane i holds layers i*LPL ..): given the composite of each lane's own layers
// (its top layer applied last), return the composite of everything BELOW the lane, i.e. of lanes i+1 .. 31 (lane 31: identity).
template <class M, class MulT>
__device__ __forceinline__ M suffix_exclusive(M own, int lane, M identity, MulT mul) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const M q = shfl_down(own, d);
    if (lane + d < 32) own = mul(own, q);   // own covers lanes [i, i+d), q covers [i+d, i+2d): q is applied first
  }
  M below = shfl_down(own, 1);
  if (lane == 31) below = identity;
  return below;
}
// Downward scans: composite of each lane's own layers (its bottom layer applied last) -> composite of everything ABOVE the lane
// (lanes 0 .. i-1; lane 0: identity).
template <class M, class MulT>
__device__ __forceinline__ M prefix_exclusive(M own, int lane, M identity, MulT mul) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const M q = shfl_up(own, d);
    if (lane >= d) own = mul(own, q);       // own covers lanes (i-d, i], q covers (i-2d, i-d]: q is applied first
  }
  M above = shfl_up(own, 1);
  if (lane == 0) above = identity;
  return above;
}

}  // namespace ecb
