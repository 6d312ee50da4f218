// qg_math.cuh -- register-resident 3-vector / 3x3 helpers for the per-lane rigid-body code.
#pragma once
#include <cuda_runtime.h>

#define DI __device__ __forceinline__

struct v3 { float x, y, z; };
struct m3 { v3 r0, r1, r2; };                 // rows
struct s3 { float xx, yy, zz, xy, xz, yz; };  // symmetric

DI v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
DI v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
DI v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
DI v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
DI v3 operator*(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
DI void operator+=(v3& a, v3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
DI void operator-=(v3& a, v3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
DI float dot(v3 a, v3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
DI v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DI v3 fma3(float s, v3 a, v3 b) { return V3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }  // s*a + b
DI v3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }

DI v3 mul(const m3& A, v3 v) { return V3(dot(A.r0, v), dot(A.r1, v), dot(A.r2, v)); }
DI v3 tmul(const m3& A, v3 v) { return fma3(v.x, A.r0, fma3(v.y, A.r1, v.z * A.r2)); }  // A^T v
DI v3 col0(const m3& A) { return V3(A.r0.x, A.r1.x, A.r2.x); }
DI v3 col1(const m3& A) { return V3(A.r0.y, A.r1.y, A.r2.y); }
DI v3 col2(const m3& A) { return V3(A.r0.z, A.r1.z, A.r2.z); }
DI m3 ldm3(const float* p) { m3 A; A.r0 = ld3(p); A.r1 = ld3(p + 3); A.r2 = ld3(p + 6); return A; }
DI m3 matmul(const m3& A, const m3& B) {
    m3 C;
    C.r0 = fma3(A.r0.x, B.r0, fma3(A.r0.y, B.r1, A.r0.z * B.r2));
    C.r1 = fma3(A.r1.x, B.r0, fma3(A.r1.y, B.r1, A.r1.z * B.r2));
    C.r2 = fma3(A.r2.x, B.r0, fma3(A.r2.y, B.r1, A.r2.z * B.r2));
    return C;
}
DI v3 mul(const s3& S, v3 v) {
    return V3(fmaf(S.xx, v.x, fmaf(S.xy, v.y, S.xz * v.z)), fmaf(S.xy, v.x, fmaf(S.yy, v.y, S.yz * v.z)),
              fmaf(S.xz, v.x, fmaf(S.yz, v.y, S.zz * v.z)));
}
// R * S * R^T for symmetric S
DI s3 rot_sym(const m3& R, const s3& S) {
    v3 t0 = mul(S, R.r0), t1 = mul(S, R.r1), t2 = mul(S, R.r2);  // S * (row_i of R)^T
    s3 o;
    o.xx = dot(R.r0, t0); o.yy = dot(R.r1, t1); o.zz = dot(R.r2, t2);
    o.xy = dot(R.r0, t1); o.xz = dot(R.r0, t2); o.yz = dot(R.r1, t2);
    return o;
}

// integer quad sum by xor-shuffles (used a handful of times per launch; the hot float reductions go through QuadRed)
DI int qsumi(int v, unsigned qm) {
    v += __shfl_xor_sync(qm, v, 1);
    v += __shfl_xor_sync(qm, v, 2);
    return v;
}

// Quad all-reduce through shared memory: every lane stores its partials ([slot][lane] layout, one 32-float row
// per slot and warp), one __syncwarp over the quad, then each lane reads the 4 partials of its quad with a single
// 128-bit load.  All four lanes add the same numbers in the same order, so the sums are bit-identical across the
// quad (control flow derived from them stays quad-uniform).  Compared with xor-shuffles under a non-constant
// member mask (WARPSYNC + SHFL + reconvergence scaffolding per value) this is ~3x fewer instructions.
#define QG_QR_SLOTS 28
struct QuadRed {
    float* s;      // this warp's scratch: QG_QR_SLOTS rows of 32 floats
    int lane;
    unsigned qm;
};
DI void qr_put(const QuadRed& q, int slot, float v) { q.s[slot * 32 + q.lane] = v; }
DI void qr_sync(const QuadRed& q) { __syncwarp(q.qm); }
DI float qr_get(const QuadRed& q, int slot) {
    float4 t = *reinterpret_cast<const float4*>(q.s + slot * 32 + (q.lane & ~3));
    return (t.x + t.y) + (t.z + t.w);
}

// lower-triangular packed index of a symmetric 6x6, i >= j
#define IX6(i, j) ((i) * ((i) + 1) / 2 + (j))
