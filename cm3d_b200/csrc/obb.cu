// KITTI orientation: open3d's oriented bounding box of an instance's points and the yaw the reference
// derives from it (src/kitti/2d_to_3d.py:855-876 `get_depth_bbox`, :1481-1484 fallback, :1524
// `as_euler('zyx')[0]`).
//
// The reference calls open3d 0.15.2 `get_oriented_bounding_box()` (un-vendored pip dependency, absent
// here).  Its published algorithm (OrientedBoundingBox::CreateFromPoints) is restated in
// oracle/obb_oracle.py and implemented here:
//   convex hull (Qhull)  ->  k_hull_obb: per instance, the set of HULL VERTICES by gift wrapping
//   mean, covariance of the hull vertices (E[xx^T] - E[x]E[x]^T, fp64) -> symmetric eigen-decomposition
//   (cyclic Jacobi, fp64) -> columns sorted by descending eigenvalue, col2 = col0 x col1
//   -> extent / centre from the min/max of R^T (v - mean) over the hull vertices
//   -> the reference's axis shuffle: axes sorted by axis-aligned size ascending,
//      wlh = [extent[idx x], extent[idx y], extent[idx z]], R' = [R[:,idx z], R[:,idx y], R[:,idx x]]
//   -> yaw = scipy 1.11.4 `Rotation.from_matrix(R').as_euler('zyx')[0]` restated (quaternion by the
//      largest of m00/m11/m22/trace, no determinant check - R' is left-handed for odd shuffles)
// PARITY UNPINNED against the open3d binary in one respect: the SIGN of Eigen's eigenvectors is
// implementation defined; here (and in the oracle) each of the first two columns is flipped so that its
// largest-magnitude component is positive.  A cloud Qhull rejects (flat: collinear / coplanar points)
// makes the reference's bare `except` substitute centre = first point, extent 1, R = identity; the
// kernel does the same when its hull is flat (exact test) or its wrapping does not close.
//
// Gift wrapping (one block of 1024 threads per instance, fp64 on the fp32 coordinates, differences
// against the edge origin are exact): start from the lexicographically smallest point a, the next
// point b of the 2-D hull of the xy projection (so the vertical plane through ab supports the cloud),
// and wrap around directed edges: the face (s,t,p) across edge (s,t) is the p for which no point q
// has det[t-s, p-s, q-s] > 0 - a tournament (thread-local, then warp shuffles, then across warps)
// whose pairwise comparison is always evaluated with the lower point index first, so every lane
// reaches the same decision.  Faces go to a per-instance list in the workspace; a directed edge is
// open while no listed face contains it (block-wide scan of the list, a few hundred to a few thousand
// faces).  O(faces x points) determinant evaluations per instance: h ~ 10^2..10^3 of M <= ~2*10^4.
// Exact ties (four coplanar hull points, duplicates) take the candidate on the far side of the old
// face, then the one farther from the edge, then the lower index; Qhull's own tolerance-based facet
// merging is NOT reproduced (it only matters for points within ~1e-13 m of a facet plane).
#include "common.cuh"

namespace cm3d {

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane_id() == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}
__device__ __forceinline__ float block_min(float v, float *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane_id() == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fminf(t, s_red[w]);
    return t;
}

// cyclic Jacobi on a symmetric 3x3 (a), eigenvectors in the columns of v
__device__ void jacobi3(double a[3][3], double v[3][3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {       // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {       // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------ hull vertices
// fp64 copies of the wrapped points in shared memory were tried (no float -> double conversion inside the tournaments):
// they cost a quarter of the survivor capacity and ran 2.4x slower on C3 (6.2 ms against 2.6 ms per 32 frames)
#ifndef CM3D_HULL_DOUBLE
#define CM3D_HULL_DOUBLE 0
#endif
#ifndef CM3D_HULL_CACHED
#define CM3D_HULL_CACHED 1
#endif
#if CM3D_HULL_DOUBLE
typedef double hull_coord_t;
#else
typedef float hull_coord_t;
#endif
constexpr int kHullThreads = 512;

struct HullPts {
    const float *x, *y, *z;            // fp32 coordinates in global memory, or ...
    const hull_coord_t *dx, *dy, *dz;  // ... copies in shared memory (dx != nullptr; fp64: no conversion inside the tournaments)
    int m;
    __device__ __forceinline__ void get(int r, double &px, double &py, double &pz) const
    {
        if (dx) { px = (double)dx[r]; py = (double)dy[r]; pz = (double)dz[r]; }
        else { px = (double)x[r]; py = (double)y[r]; pz = (double)z[r]; }
    }
};

// Directed edge being wrapped: origin S, direction e = T - S, and the third vertex of the face the
// edge came from, relative to S (wz; the tie rule for coplanar neighbours needs its side).
struct HullEdge {
    double sx, sy, sz, ex, ey, ez, zx, zy, zz;
};

// Should candidate b replace candidate a as third vertex of the face on edge E?  Evaluated on
// (lower index, higher index) whatever the argument order, so both sides of a shuffle agree.
__device__ bool hull_beats(const HullPts &P, const HullEdge &E, int a, int b)
{
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    double lx, ly, lz, hx, hy, hz;
    P.get(lo, lx, ly, lz);
    P.get(hi, hx, hy, hz);
    if (lx == hx && ly == hy && lz == hz) return b == lo;     // duplicate points: always the lower index (one copy per vertex)
    lx -= E.sx; ly -= E.sy; lz -= E.sz;
    hx -= E.sx; hy -= E.sy; hz -= E.sz;
    const double nlx = E.ey * lz - E.ez * ly, nly = E.ez * lx - E.ex * lz, nlz = E.ex * ly - E.ey * lx;   // e x w_lo
    const double d = nlx * hx + nly * hy + nlz * hz;          // > 0: hi lies outside the plane (S, T, lo)
    if (d > 0.0) return b == hi;
    if (d < 0.0) return b == lo;
    const double nhx = E.ey * hz - E.ez * hy, nhy = E.ez * hx - E.ex * hz, nhz = E.ex * hy - E.ey * hx;
    const double nnl = nlx * nlx + nly * nly + nlz * nlz, nnh = nhx * nhx + nhy * nhy + nhz * nhz;
    if (nnh == 0.0) return b == lo && nnl > 0.0;              // a point on the edge's line is never a third vertex
    if (nnl == 0.0) return b == hi;
    if (nlx * nhx + nly * nhy + nlz * nhz < 0.0) {            // coplanar, on opposite sides of the edge:
        const double nzx = E.ey * E.zz - E.ez * E.zy, nzy = E.ez * E.zx - E.ex * E.zz, nzz = E.ex * E.zy - E.ey * E.zx;
        const double side_lo = nlx * nzx + nly * nzy + nlz * nzz;       // the far side of the old face wins
        return side_lo < 0.0 ? b == lo : b == hi;
    }
    if (nnh > nnl) return b == hi;                             // same side: farther from the edge, then lower index
    return b == lo;
}

// p such that no point lies outside the plane (S, T, p); -1 when every point is on the edge's line.
__device__ int hull_wrap(const HullPts &P, const HullEdge &E, int s, int t, int *s_cand)
{
    int best = -1;
    for (int r = threadIdx.x; r < P.m; r += blockDim.x) {
        if (r == s || r == t) continue;
        if (best < 0) {
            double wx, wy, wz;
            P.get(r, wx, wy, wz);
            wx -= E.sx; wy -= E.sy; wz -= E.sz;
            const double nx = E.ey * wz - E.ez * wy, ny = E.ez * wx - E.ex * wz, nz = E.ex * wy - E.ey * wx;
            if (nx * nx + ny * ny + nz * nz > 0.0) best = r;
        } else if (hull_beats(P, E, best, r)) {
            best = r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other >= 0 && other != best && (best < 0 || hull_beats(P, E, best, other))) best = other;
    }
    __syncthreads();
    if (lane_id() == 0) s_cand[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int b2 = threadIdx.x < (blockDim.x >> 5) ? s_cand[threadIdx.x] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int other = __shfl_xor_sync(0xffffffffu, b2, o);
            if (other >= 0 && other != b2 && (b2 < 0 || hull_beats(P, E, b2, other))) b2 = other;
        }
        if (threadIdx.x == 0) s_cand[32] = b2;
    }
    __syncthreads();
    return s_cand[32];
}

// lexicographic (x, y, z, index) minimum over the block
__device__ int hull_lexmin(const HullPts &P, int *s_cand)
{
    auto less = [&](int a, int b) {
        double ax, ay, az, bx, by, bz;
        P.get(a, ax, ay, az);
        P.get(b, bx, by, bz);
        if (ax != bx) return ax < bx;
        if (ay != by) return ay < by;
        if (az != bz) return az < bz;
        return a < b;
    };
    int best = -1;
    for (int r = threadIdx.x; r < P.m; r += blockDim.x)
        if (best < 0 || less(r, best)) best = r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other >= 0 && (best < 0 || less(other, best))) best = other;
    }
    __syncthreads();
    if (lane_id() == 0) s_cand[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b2 = -1;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            if (s_cand[w] >= 0 && (b2 < 0 || less(s_cand[w], b2))) b2 = s_cand[w];
        s_cand[32] = b2;
    }
    __syncthreads();
    return s_cand[32];
}

// next vertex after a on the counter-clockwise 2-D hull of the xy projection: every other point is
// to the left of a -> b (ties: the farther one, then the lower index); -1 when all share a's (x, y).
__device__ int hull_next2d(const HullPts &P, int a, int *s_cand)
{
    double ax, ay, az;
    P.get(a, ax, ay, az);
    auto beats = [&](int p, int q) {       // should q replace p?
        const int lo = p < q ? p : q, hi = p < q ? q : p;
        double lx, ly, lz, hx, hy, hz;
        P.get(lo, lx, ly, lz);
        P.get(hi, hx, hy, hz);
        lx -= ax; ly -= ay; hx -= ax; hy -= ay;
        const double c = lx * hy - ly * hx;            // < 0: hi is to the right of a -> lo
        if (c < 0.0) return q == hi;
        if (c > 0.0) return q == lo;
        const double dl = lx * lx + ly * ly, dh = hx * hx + hy * hy;
        if (dh > dl) return q == hi;
        return q == lo;
    };
    auto valid = [&](int r) {
        double x, y, z;
        P.get(r, x, y, z);
        return r != a && (x != ax || y != ay);
    };
    int best = -1;
    for (int r = threadIdx.x; r < P.m; r += blockDim.x) {
        if (!valid(r)) continue;
        if (best < 0 || beats(best, r)) best = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other >= 0 && other != best && (best < 0 || beats(best, other))) best = other;
    }
    __syncthreads();
    if (lane_id() == 0) s_cand[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b2 = -1;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            if (s_cand[w] >= 0 && s_cand[w] != b2 && (b2 < 0 || beats(b2, s_cand[w]))) b2 = s_cand[w];
        s_cand[32] = b2;
    }
    __syncthreads();
    return s_cand[32];
}

// is the directed edge (s, t) part of a listed face?
__device__ bool hull_edge_listed(const int32_t *faces, int n_faces, int s, int t)
{
    int found = 0;
    for (int f = threadIdx.x; f < n_faces; f += blockDim.x) {
        const int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
        found |= (a == s && b == t) || (b == s && c == t) || (c == s && a == t);
    }
    return __syncthreads_or(found) != 0;
}

// First face of the hull: (b, a, p0) with a the lexicographic minimum, b the next vertex of the 2-D hull of the xy
// projection, p0 across the reversed edge of the virtual vertical face.  False when the cloud is flat.
__device__ bool hull_seed(const HullPts &P, int *s_cand, int &a, int &b, int &p0)
{
    a = hull_lexmin(P, s_cand);
    b = hull_next2d(P, a, s_cand);
    if (b < 0) return false;
    double ax, ay, az, bx, by, bz;
    P.get(a, ax, ay, az);
    P.get(b, bx, by, bz);
    HullEdge E;
    E.sx = bx; E.sy = by; E.sz = bz;
    E.ex = ax - bx; E.ey = ay - by; E.ez = az - bz;
    E.zx = E.ex; E.zy = E.ey; E.zz = E.ez + 1.0;
    p0 = hull_wrap(P, E, b, a, s_cand);
    if (p0 < 0) return false;
    // flat cloud: every point in the plane of the first face (exact test)
    double px, py, pz;
    P.get(p0, px, py, pz);
    px -= bx; py -= by; pz -= bz;
    const double nx = E.ey * pz - E.ez * py, ny = E.ez * px - E.ex * pz, nz = E.ex * py - E.ey * px;
    int off = 0;
    for (int r = threadIdx.x; r < P.m; r += blockDim.x) {
        double qx, qy, qz;
        P.get(r, qx, qy, qz);
        off |= (nx * (qx - bx) + ny * (qy - by) + nz * (qz - bz)) != 0.0;
    }
    return __syncthreads_or(off) != 0;
}

// Flags the convex-hull vertices of the block's instance in vflag[0..m), one edge at a time with the whole block
// (the fallback for clouds whose survivors do not fit shared memory).  Returns the number of faces
// (>= 4), or 0 when the cloud is flat (collinear / coplanar) or the wrapping did not close.
__device__ int hull_vertices(const HullPts &P, int32_t *faces, int face_cap, uint8_t *vflag, int *s_cand)
{
    for (int r = threadIdx.x; r < P.m; r += blockDim.x) vflag[r] = 0;
    int a, b, p0;
    if (!hull_seed(P, s_cand, a, b, p0)) return 0;
    HullEdge E;
    int n_faces = 1;
    if (threadIdx.x == 0) {
        faces[0] = b; faces[1] = a; faces[2] = p0;
        vflag[a] = vflag[b] = vflag[p0] = 1;
    }
    __syncthreads();
    for (int k = 0; k < n_faces; ++k) {
        const int fx = faces[3 * k], fy = faces[3 * k + 1], fz = faces[3 * k + 2];
        for (int j = 0; j < 3; ++j) {
            const int u = j == 0 ? fx : (j == 1 ? fy : fz);
            const int v = j == 0 ? fy : (j == 1 ? fz : fx);
            const int w = j == 0 ? fz : (j == 1 ? fx : fy);
            if (hull_edge_listed(faces, n_faces, v, u)) continue;          // the neighbour across (u, v) exists
            double vx, vy, vz, ux, uy, uz, wx, wy, wz;
            P.get(v, vx, vy, vz);
            P.get(u, ux, uy, uz);
            P.get(w, wx, wy, wz);
            E.sx = vx; E.sy = vy; E.sz = vz;
            E.ex = ux - vx; E.ey = uy - vy; E.ez = uz - vz;
            E.zx = wx - vx; E.zy = wy - vy; E.zz = wz - vz;
            const int p = hull_wrap(P, E, v, u, s_cand);
            if (p < 0 || p == w || n_faces >= face_cap) return 0;
            if (threadIdx.x == 0) {
                faces[3 * n_faces] = v; faces[3 * n_faces + 1] = u; faces[3 * n_faces + 2] = p;
                vflag[p] = 1;
            }
            ++n_faces;
            __syncthreads();
        }
    }
    return n_faces;
}

// ------------------------------------------------------------------------------------------ one edge per WARP
// The block-wide wrap above pays ~8 barriers and two reduction levels per face, and a hull has 10^2..10^3 faces.
// With the points in shared memory every warp wraps its OWN open edge: the open edges sit in a queue, the
// directed edges of the faces found so far in a hash set, both in shared memory.  A face is committed through
// its canonical directed edge (the rotation that starts at its smallest vertex): whichever warp inserts that
// edge first owns the face, so a face reached from two of its edges at the same time is counted once.
constexpr int kEdgeTab = 8192;               // directed-edge hash set (open addressing), entries: key + 1
constexpr int kEdgeMax = kEdgeTab * 3 / 4;
constexpr int kQueueCap = 8192;

struct WrapTables {
    uint32_t etab[kEdgeTab];
    uint32_t q_st[kQueueCap];                // (s << 16 | t) + 1, 0 = not yet written
    uint16_t q_w[kQueueCap];                 // third vertex of the face the open edge borders
    int tail, head, outstanding, n_faces, n_edges, fail;
};

__device__ __forceinline__ uint32_t edge_hash(uint32_t key) { return (key * 2654435761u) >> 19; }      // 13 bits

// true when the directed edge was not in the set before
__device__ bool edge_insert(WrapTables &W, uint32_t key)
{
    uint32_t h = edge_hash(key) & (kEdgeTab - 1);
    for (int probe = 0; probe < kEdgeTab; ++probe) {
        const uint32_t old = atomicCAS(&W.etab[h], 0u, key + 1u);
        if (old == 0u) {
            if (atomicAdd(&W.n_edges, 1) >= kEdgeMax) atomicExch(&W.fail, 1);
            return true;
        }
        if (old == key + 1u) return false;
        h = (h + 1) & (kEdgeTab - 1);
    }
    atomicExch(&W.fail, 1);
    return false;
}

__device__ bool edge_present(WrapTables &W, uint32_t key)
{
    uint32_t h = edge_hash(key) & (kEdgeTab - 1);
    for (int probe = 0; probe < kEdgeTab; ++probe) {
        const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&W.etab[h]);
        if (cur == 0u) return false;
        if (cur == key + 1u) return true;
        h = (h + 1) & (kEdgeTab - 1);
    }
    return false;
}

__device__ void queue_push(WrapTables &W, int s, int t, int w)
{
    atomicAdd(&W.outstanding, 1);
    const int i = atomicAdd(&W.tail, 1);
    if (i >= kQueueCap) { atomicExch(&W.fail, 1); atomicSub(&W.outstanding, 1); return; }
    W.q_w[i] = (uint16_t)w;
    __threadfence_block();
    *reinterpret_cast<volatile uint32_t *>(&W.q_st[i]) = (((uint32_t)s << 16) | (uint32_t)t) + 1u;
}

// lane 0 only.  Commits face (s, t, p) unless it exists; pushes its still-open neighbours' edges.
__device__ void face_commit(WrapTables &W, int s, int t, int p, uint8_t *vflag, int32_t *faces, int face_cap)
{
    int a = s, b = t, c = p;
    if (b < a && b < c) { a = t; b = p; c = s; }
    else if (c < a && c < b) { a = p; b = s; c = t; }
    if (!edge_insert(W, ((uint32_t)a << 16) | (uint32_t)b)) return;           // another warp owns this face
    const bool ok1 = edge_insert(W, ((uint32_t)b << 16) | (uint32_t)c);
    const bool ok2 = edge_insert(W, ((uint32_t)c << 16) | (uint32_t)a);
    if (!ok1 || !ok2) atomicExch(&W.fail, 2);                                 // a directed edge in two faces: not a hull
    const int f = atomicAdd(&W.n_faces, 1);
    if (faces) {
        if (f < face_cap) { faces[3 * f] = a; faces[3 * f + 1] = b; faces[3 * f + 2] = c; }
        else atomicExch(&W.fail, 1);
    }
    vflag[a] = 1; vflag[b] = 1; vflag[c] = 1;
    if (!edge_present(W, ((uint32_t)b << 16) | (uint32_t)a)) queue_push(W, b, a, c);
    if (!edge_present(W, ((uint32_t)c << 16) | (uint32_t)b)) queue_push(W, c, b, a);
    if (!edge_present(W, ((uint32_t)a << 16) | (uint32_t)c)) queue_push(W, a, c, b);
}

// p such that no point lies outside the plane (S, T, p), found by ONE warp (-1: every point on the edge's line)
__device__ int hull_wrap_warp(const HullPts &P, const HullEdge &E, int s, int t)
{
    const int lane = (int)lane_id();
    // thread-local pass with the normal of the current best plane cached: one difference and one dot product per
    // point; the full comparison (tie rules, lower index first) only when the dot product is exactly zero
    int best = -1;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    for (int r = lane; r < P.m; r += 32) {
        if (r == s || r == t) continue;
        double wx, wy, wz;
        P.get(r, wx, wy, wz);
        wx -= E.sx; wy -= E.sy; wz -= E.sz;
        bool take;
        if (best < 0) {
            take = true;
        } else {
#if CM3D_HULL_CACHED
            const double d = nx * wx + ny * wy + nz * wz;         // > 0: r lies outside the plane (S, T, best)
            take = d > 0.0 || (d == 0.0 && hull_beats(P, E, best, r));
#else
            take = hull_beats(P, E, best, r);
#endif
        }
        if (take) {
            const double cx = E.ey * wz - E.ez * wy, cy = E.ez * wx - E.ex * wz, cz = E.ex * wy - E.ey * wx;
            if (best < 0 && !(cx * cx + cy * cy + cz * cz > 0.0)) continue;      // on the edge's line: never a third vertex
            best = r; nx = cx; ny = cy; nz = cz;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other >= 0 && other != best && (best < 0 || hull_beats(P, E, best, other))) best = other;
    }
    return best;
}

// The hull of a point set that lives in shared memory (P.m <= 65535 local indices), all warps of the block
// wrapping edges at once.  Flags the vertices in vflag[0..P.m); faces (optional) receives the face list.
// Returns the number of faces, 0 for a flat cloud, -1 when the tables overflowed or the wrapping was
// inconsistent (the caller falls back to the block-wide version).
__device__ int hull_vertices_par(const HullPts &P, WrapTables &W, int32_t *faces, int face_cap, uint8_t *vflag, int *s_cand)
{
    for (int r = threadIdx.x; r < P.m; r += blockDim.x) vflag[r] = 0;
    for (int k = threadIdx.x; k < kEdgeTab; k += blockDim.x) W.etab[k] = 0u;
    for (int k = threadIdx.x; k < kQueueCap; k += blockDim.x) W.q_st[k] = 0u;
    if (threadIdx.x == 0) { W.tail = W.head = W.outstanding = W.n_faces = W.n_edges = W.fail = 0; }
    int a, b, p0;
    if (!hull_seed(P, s_cand, a, b, p0)) return 0;            // (barriers inside: the tables are clear past this point)
    if (threadIdx.x == 0) face_commit(W, b, a, p0, vflag, faces, face_cap);
    __syncthreads();
    const int lane = (int)lane_id();
    for (int iter = 0;; ++iter) {
        // lane 0 takes the next open edge: slot index, -1 = nothing queued right now, -2 = all done (or failed)
        int slot = -1;
        if (iter > (1 << 21) && lane == 0) atomicExch(&W.fail, 1);            // never spin forever
        if (lane == 0) {
            if (*reinterpret_cast<volatile int *>(&W.fail)) slot = -2;
            else {
                const int h = *reinterpret_cast<volatile int *>(&W.head);
                const int tl = min(*reinterpret_cast<volatile int *>(&W.tail), kQueueCap);
                if (h < tl) {
                    if (atomicCAS(&W.head, h, h + 1) == h) slot = h;
                } else if (*reinterpret_cast<volatile int *>(&W.outstanding) == 0) slot = -2;
            }
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot == -2) break;
        if (slot == -1) { __nanosleep(100); continue; }
        uint32_t key = 0;
        int w = 0;
        if (lane == 0) {
            int spin = 0;
            while ((key = *reinterpret_cast<volatile uint32_t *>(&W.q_st[slot])) == 0u && ++spin < (1 << 24)) {}     // the pusher writes it last
            if (key == 0u) { atomicExch(&W.fail, 1); key = 1u; }
            key -= 1u;
            w = (int)*reinterpret_cast<volatile uint16_t *>(&W.q_w[slot]);
            if (edge_present(W, key)) key = 0xffffffffu;                                            // closed meanwhile
        }
        key = __shfl_sync(0xffffffffu, key, 0);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (key != 0xffffffffu) {
            const int es = (int)(key >> 16), et = (int)(key & 0xffffu);
            double sx, sy, sz, tx, ty, tz, wx, wy, wz;
            P.get(es, sx, sy, sz);
            P.get(et, tx, ty, tz);
            P.get(w, wx, wy, wz);
            HullEdge E;
            E.sx = sx; E.sy = sy; E.sz = sz;
            E.ex = tx - sx; E.ey = ty - sy; E.ez = tz - sz;
            E.zx = wx - sx; E.zy = wy - sy; E.zz = wz - sz;
            const int p = hull_wrap_warp(P, E, es, et);
            if (lane == 0) {
                if (p < 0 || p == w) atomicExch(&W.fail, 3);
                else face_commit(W, es, et, p, vflag, faces, face_cap);
            }
        }
        if (lane == 0) { __threadfence_block(); atomicSub(&W.outstanding, 1); }
        __syncwarp();
    }
    __syncthreads();
    if (W.fail) return W.fail == 3 ? 0 : -1;          // 3: a wrap came back flat (degenerate input) -> the reference's fallback box
    return W.n_faces;
}

// ------------------------------------------------------------------------------------------ prefilter
// Gift wrapping costs (faces x points) determinant evaluations, and a LiDAR segment has 10^3..10^4 points of
// which a few per cent are hull vertices.  So first: the extreme points of the cloud along up to 256 fixed
// directions (a Fibonacci sphere; fp32 dot products - ANY subset of the input works), their hull by the same
// gift wrapping (<= 256 points: cheap), and every point strictly inside that inner polytope (fp64 plane
// tests with a 2^-40 relative guard band) is dropped: it cannot be a vertex of the full hull.  ~10 % of a
// segment survives; the survivors' coordinates move to shared memory and are wrapped there.  A cloud
// whose coarse hull is flat, whose coarse face count is inconsistent, or whose survivors do not fit falls
// back to wrapping every point from global memory.
constexpr int kHullDirs = 256;
constexpr int kHullSurvCap = CM3D_HULL_DOUBLE ? 3072 : 4096;
constexpr int kHullDirect = 768;             // clouds this small are wrapped directly (in shared memory)
constexpr int kCoarseFaceCap = 2 * kHullDirs;

struct HullShared {
    hull_coord_t sx[kHullSurvCap], sy[kHullSurvCap], sz[kHullSurvCap];
    int32_t sorig[kHullSurvCap];
    uint8_t sflag[kHullSurvCap];
    double plane[kCoarseFaceCap][4];
    float4 plane32[kCoarseFaceCap];          // the same planes in fp32 (offset without the guard band) ...
    float margin32[kCoarseFaceCap];          // ... and the band inside which the fp32 value decides nothing
    int32_t cfaces[3 * kCoarseFaceCap];
    float dirs[kHullDirs][3];
    int32_t ext[kHullDirs];
    int32_t wsum[kHullThreads / 32];
    int32_t count;
    WrapTables W;
};

// Loads the whole cloud (m <= kHullSurvCap) into shared memory, identity index map.
__device__ void hull_load_all(HullShared &S, const float *sx, const float *sy, const float *sz, int m)
{
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        S.sx[r] = sx[r]; S.sy[r] = sy[r]; S.sz[r] = sz[r]; S.sorig[r] = r;
    }
    __syncthreads();
}

// Extreme points along n_dir directions -> S.ext[0..n_dir)
__device__ void hull_extremes(HullShared &S, const float *sx, const float *sy, const float *sz, int m, int n_dir)
{
    __shared__ float s_bv[kHullThreads / 32][8];
    __shared__ int s_bi[kHullThreads / 32][8];
    for (int k = threadIdx.x; k < n_dir; k += blockDim.x) {
        const float c = 1.0f - 2.0f * ((float)k + 0.5f) / (float)n_dir;
        const float sn = sqrtf(fmaxf(1.0f - c * c, 0.0f));
        float st, ct;
        sincosf(2.39996322972865332f * (float)k, &st, &ct);
        S.dirs[k][0] = ct * sn; S.dirs[k][1] = st * sn; S.dirs[k][2] = c;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = (int)lane_id();
    for (int d0 = 0; d0 < n_dir; d0 += 8) {
        float bv[8];
        int bi[8];
        float dx[8], dy[8], dz[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bv[k] = -INFINITY; bi[k] = -1;
            const int d = min(d0 + k, n_dir - 1);
            dx[k] = S.dirs[d][0]; dy[k] = S.dirs[d][1]; dz[k] = S.dirs[d][2];
        }
        for (int r = threadIdx.x; r < m; r += blockDim.x) {
            const float x = sx[r], y = sy[r], z = sz[r];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float v = dx[k] * x + dy[k] * y + dz[k] * z;
                if (v > bv[k]) { bv[k] = v; bi[k] = r; }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv[k], o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi[k], o);
                if (ov > bv[k] || (ov == bv[k] && oi >= 0 && (bi[k] < 0 || oi < bi[k]))) { bv[k] = ov; bi[k] = oi; }
            }
            if (lane == 0) { s_bv[warp][k] = bv[k]; s_bi[warp][k] = bi[k]; }
        }
        __syncthreads();
        if (threadIdx.x < 8 && d0 + (int)threadIdx.x < n_dir) {
            float v = -INFINITY;
            int idx = -1;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                if (s_bv[w][threadIdx.x] > v || (s_bv[w][threadIdx.x] == v && s_bi[w][threadIdx.x] >= 0 && (idx < 0 || s_bi[w][threadIdx.x] < idx))) {
                    v = s_bv[w][threadIdx.x]; idx = s_bi[w][threadIdx.x];
                }
            S.ext[d0 + threadIdx.x] = idx;
        }
        __syncthreads();
    }
}

// Flags of the hull vertices into vflag[0..m).  Returns (faces, vertices) of the hull, faces = 0 when flat.
__device__ int hull_vertices_filtered(HullShared &S, const float *sx, const float *sy, const float *sz, int m,
                                      int32_t *faces_ws, uint8_t *vflag, int *s_cand)
{
    for (int r = threadIdx.x; r < m; r += blockDim.x) vflag[r] = 0;
    int ns = -1;                                  // survivors in shared memory (-1: wrap every point from global memory)
    if (m <= kHullDirect) {
        hull_load_all(S, sx, sy, sz, m);
        ns = m;
    } else {
        int n_dir = min(kHullDirs, max(32, m / 32)) & ~7;
        hull_extremes(S, sx, sy, sz, m, n_dir);
        // distinct extreme points, in direction order -> the coarse set in S.sx/sy/sz[0..nc)
        {   // thread k owns direction k: unique when no earlier direction found the same point; ranks by ballot
            const int k = threadIdx.x;
            const int e = k < n_dir ? S.ext[k] : -1;
            bool uniq = e >= 0;
            for (int j = 0; j < k && uniq; ++j) uniq = S.ext[j] != e;
            const unsigned bal = __ballot_sync(0xffffffffu, uniq);
            if (lane_id() == 0) S.wsum[threadIdx.x >> 5] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { if (w < (int)(threadIdx.x >> 5)) before += S.wsum[w]; total += S.wsum[w]; }
            if (uniq) {
                const int at = before + __popc(bal & lanemask_lt());
                S.sorig[at] = e; S.sx[at] = sx[e]; S.sy[at] = sy[e]; S.sz[at] = sz[e];
            }
            if (threadIdx.x == 0) S.count = total;
            __syncthreads();
        }
        const int nc = S.count;
        HullPts C{nullptr, nullptr, nullptr, S.sx, S.sy, S.sz, nc};
        const int nf = nc >= 4 ? hull_vertices_par(C, S.W, S.cfaces, kCoarseFaceCap, S.sflag, s_cand) : 0;
        __syncthreads();
        bool usable = nf >= 4;
        if (usable) {                             // Euler check of the coarse hull, then its planes
            int v = 0;
            for (int r = threadIdx.x; r < nc; r += blockDim.x) v += S.sflag[r];
            int tot = 0;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
            if (lane_id() == 0) S.wsum[threadIdx.x >> 5] = v;
            __syncthreads();
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += S.wsum[w];
            usable = nf == 2 * tot - 4;
            __syncthreads();
        }
        if (usable) {
            // |p|_1 of the farthest point: the scale of every plane evaluation's rounding error
            float r1 = 0.0f;
            for (int r = threadIdx.x; r < m; r += blockDim.x) r1 = fmaxf(r1, fabsf(sx[r]) + fabsf(sy[r]) + fabsf(sz[r]));
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, o2));
            if (lane_id() == 0) S.wsum[threadIdx.x >> 5] = __float_as_int(r1);
            __syncthreads();
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r1 = fmaxf(r1, __int_as_float(S.wsum[w]));
            const double Rcloud = (double)r1 * 1.000001;
            __syncthreads();
            for (int f = threadIdx.x; f < nf; f += blockDim.x) {
                const int a = S.cfaces[3 * f], b = S.cfaces[3 * f + 1], c = S.cfaces[3 * f + 2];
                const double ax = S.sx[a], ay = S.sy[a], az = S.sz[a];
                const double ux = (double)S.sx[b] - ax, uy = (double)S.sy[b] - ay, uz = (double)S.sz[b] - az;
                const double vx = (double)S.sx[c] - ax, vy = (double)S.sy[c] - ay, vz = (double)S.sz[c] - az;
                const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;     // outward
                const double off = -(nx * ax + ny * ay + nz * az);
                // guard band: |rounding of n.p + off| <= ~8 ulp of |n|_1 * max|coordinate|; 2^-40 is 10^3 times that
                const double R = Rcloud + fabs(ax) + fabs(ay) + fabs(az);      // |n.p| + |off| <= |n|_1 R
                S.plane[f][0] = nx; S.plane[f][1] = ny; S.plane[f][2] = nz;
                S.plane[f][3] = off + (fabs(nx) + fabs(ny) + fabs(nz)) * R * 0x1p-40;                   // inside iff n.p + this < 0
                // fp32 screen of the same test: rounding the four constants and three FMAs is below 2^-21 |n|_1 R
                S.plane32[f] = make_float4((float)nx, (float)ny, (float)nz, (float)off);
                S.margin32[f] = (float)((fabs(nx) + fabs(ny) + fabs(nz)) * R * 0x1p-19);
            }
            __syncthreads();
            // survivors, in index order (ballot ranks + a scan of the warp totals per chunk of blockDim points)
            int base = 0;
            bool overflow = false;
            const int warp = threadIdx.x >> 5, lane = (int)lane_id(), nw = (int)(blockDim.x >> 5);
            for (int c0 = 0; c0 < m; c0 += blockDim.x) {
                const int r = c0 + threadIdx.x;
                bool keep = false;
                float x = 0.f, y = 0.f, z = 0.f;
                if (r < m) {
                    x = sx[r]; y = sy[r]; z = sz[r];
                    const double px = x, py = y, pz = z;
                    for (int f = 0; f < nf; ++f) {
                        const float4 pl = S.plane32[f];
                        const float v = fmaf(pl.x, x, fmaf(pl.y, y, fmaf(pl.z, z, pl.w)));
                        const float mg = S.margin32[f];
                        if (v < -mg) continue;                              // clearly inside this face
                        if (v > mg ||                                       // clearly outside; in between fp64 decides
                            !(S.plane[f][0] * px + S.plane[f][1] * py + S.plane[f][2] * pz + S.plane[f][3] < 0.0)) { keep = true; break; }
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) S.wsum[warp] = __popc(bal);
                __syncthreads();
                int before = 0, total = 0;
                for (int w = 0; w < nw; ++w) { if (w < warp) before += S.wsum[w]; total += S.wsum[w]; }
                if (base + total > kHullSurvCap) overflow = true;
                else if (keep) {
                    const int at = base + before + __popc(bal & lanemask_lt());
                    S.sx[at] = x; S.sy[at] = y; S.sz[at] = z; S.sorig[at] = r;
                }
                base += total;
                __syncthreads();
                if (overflow) break;
            }
            if (!overflow) ns = base;
        }
        if (ns < 0 && m <= kHullSurvCap) {        // no usable coarse hull, but the cloud fits: wrap it in shared memory
            __syncthreads();
            hull_load_all(S, sx, sy, sz, m);
            ns = m;
        }
    }
    __syncthreads();
    if (ns < 0) {                                 // fallback: every point, from global memory
        HullPts P{sx, sy, sz, nullptr, nullptr, nullptr, m};
        return hull_vertices(P, faces_ws, 2 * m, vflag, s_cand);
    }
    HullPts P{nullptr, nullptr, nullptr, S.sx, S.sy, S.sz, ns};
    int n_faces = hull_vertices_par(P, S.W, nullptr, 0, S.sflag, s_cand);
    __syncthreads();
    if (n_faces < 0) n_faces = hull_vertices(P, faces_ws, 2 * m, S.sflag, s_cand);       // tables overflowed: one edge at a time
    __syncthreads();
    if (n_faces > 0)
        for (int r = threadIdx.x; r < ns; r += blockDim.x)
            if (S.sflag[r]) vflag[S.sorig[r]] = 1;
    __syncthreads();
    return n_faces;
}

// ------------------------------------------------------------------------------------------ the box
// mode 0: hull vertices (open3d); mode 1: all member points (round 1's estimator, kept for comparison)
__global__ void __launch_bounds__(kHullThreads, 1)
k_hull_obb(const float *__restrict__ seg_xyzw, int64_t seg_cap, const int32_t *__restrict__ seg_off, int min_pts, int mode,
           const int32_t *__restrict__ order, int32_t *__restrict__ hull_ws, float *__restrict__ obb,
           int32_t *__restrict__ hull_info, const int32_t *__restrict__ errflags)
{
    __shared__ double s_red[32];
    __shared__ float s_redf[32];
    __shared__ int s_cand[33];
    __shared__ double s_R[3][3];
    __shared__ double s_mean[3];
    const int i = order ? order[blockIdx.x] : (int)blockIdx.x;       // largest instances first (the medoid's schedule)
    float *out = obb + (size_t)i * 16;
    const float nanv = __int_as_float(0x7fc00000);
    const int o = seg_off[i], m = seg_off[i + 1] - o;
    if (errflags[CM3D_ERR_SEG_OVERFLOW] != 0 || m < min_pts || m < 1) {
        if (threadIdx.x < 16) out[threadIdx.x] = nanv;
        if (threadIdx.x == 0 && hull_info) hull_info[i] = 0;
        return;
    }
    const float *sx = seg_xyzw + o, *sy = seg_xyzw + seg_cap + o, *sz = seg_xyzw + 2 * seg_cap + o;
    uint8_t *vflag = reinterpret_cast<uint8_t *>(hull_ws + 6 * seg_cap) + o;
    int n_faces = -1;
    if (mode == 0) {
        extern __shared__ __align__(16) unsigned char s_dyn[];
        HullShared &S = *reinterpret_cast<HullShared *>(s_dyn);
        n_faces = hull_vertices_filtered(S, sx, sy, sz, m, hull_ws + 6 * (int64_t)o, vflag, s_cand);
        __syncthreads();
        if (n_faces == 0) {       // kitti:1483-1484: bbox = [pts3d[0], [1,1,1], identity] -> yaw 0
            if (threadIdx.x == 0) {
                out[0] = 0.0f; out[1] = sx[0]; out[2] = sy[0]; out[3] = sz[0];
                out[4] = out[5] = out[6] = 1.0f;
                for (int k = 0; k < 9; ++k) out[7 + k] = (k % 4 == 0) ? 1.0f : 0.0f;
                if (hull_info) hull_info[i] = -1;
            }
            return;
        }
    }
    // pass 1: first and second moments of the selected points, axis-aligned size of ALL points (kitti:863-865)
    double a1[3] = {0, 0, 0}, a2[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const float x = sx[r], y = sy[r], z = sz[r];
        lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
        lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
        lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
        if (mode == 0 && !vflag[r]) continue;
        const double dx = x, dy = y, dz = z;
        ++cnt;
        a1[0] += dx; a1[1] += dy; a1[2] += dz;
        a2[0] += dx * dx; a2[1] += dx * dy; a2[2] += dx * dz; a2[3] += dy * dy; a2[4] += dy * dz; a2[5] += dz * dz;
    }
    const double n = block_sum((double)cnt, s_red);
    const double mx = block_sum(a1[0], s_red) / n, my = block_sum(a1[1], s_red) / n, mz = block_sum(a1[2], s_red) / n;
    double c[6];
    for (int k = 0; k < 6; ++k) c[k] = block_sum(a2[k], s_red) / n;
    c[0] -= mx * mx; c[1] -= mx * my; c[2] -= mx * mz; c[3] -= my * my; c[4] -= my * mz; c[5] -= mz * mz;
    float size[3];
    for (int k = 0; k < 3; ++k) size[k] = -block_min(-hi[k], s_redf) - block_min(lo[k], s_redf);
    if (threadIdx.x == 0) {
        if (hull_info) hull_info[i] = mode == 0 ? (n_faces == 2 * (int)n - 4 ? (int)n : -(int)n - 1) : (int)n;
        double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}}, v[3][3];
        jacobi3(a, v);
        int ord[3] = {0, 1, 2};                       // descending eigenvalue, stable
        for (int p = 0; p < 2; ++p)
            for (int q = 0; q < 2 - p; ++q)
                if (a[ord[q]][ord[q]] < a[ord[q + 1]][ord[q + 1]]) { const int t = ord[q]; ord[q] = ord[q + 1]; ord[q + 1] = t; }
        double R[3][3];
        for (int col = 0; col < 2; ++col) {
            double e[3] = {v[0][ord[col]], v[1][ord[col]], v[2][ord[col]]};
            const double nn = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
            int big = 0;
            if (fabs(e[1]) > fabs(e[big])) big = 1;
            if (fabs(e[2]) > fabs(e[big])) big = 2;
            const double sgn = e[big] < 0.0 ? -1.0 : 1.0;
            for (int k = 0; k < 3; ++k) R[k][col] = sgn * e[k] / nn;
        }
        R[0][2] = R[1][0] * R[2][1] - R[2][0] * R[1][1];
        R[1][2] = R[2][0] * R[0][1] - R[0][0] * R[2][1];
        R[2][2] = R[0][0] * R[1][1] - R[1][0] * R[0][1];
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) s_R[r][k] = R[r][k];
        s_mean[0] = mx; s_mean[1] = my; s_mean[2] = mz;
    }
    __syncthreads();
    // pass 2: extent / centre in the principal frame
    float plo[3] = {INFINITY, INFINITY, INFINITY}, phi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        if (mode == 0 && !vflag[r]) continue;
        const double dx = sx[r] - s_mean[0], dy = sy[r] - s_mean[1], dz = sz[r] - s_mean[2];
        for (int k = 0; k < 3; ++k) {
            const float q = (float)(s_R[0][k] * dx + s_R[1][k] * dy + s_R[2][k] * dz);
            plo[k] = fminf(plo[k], q); phi[k] = fmaxf(phi[k], q);
        }
    }
    float elo[3], ehi[3];
    for (int k = 0; k < 3; ++k) { elo[k] = block_min(plo[k], s_redf); ehi[k] = -block_min(-phi[k], s_redf); }
    if (threadIdx.x == 0) {
        // reference axis shuffle (kitti:862-876): names sorted by axis-aligned size, ascending, stable
        int axis[3] = {0, 1, 2};                      // 0='x', 1='y', 2='z'
        for (int p = 0; p < 2; ++p)
            for (int q = 0; q < 2 - p; ++q) {
                const bool gt = size[axis[q]] > size[axis[q + 1]] ||
                                (size[axis[q]] == size[axis[q + 1]] && axis[q] > axis[q + 1]);   // tuple sort: (size, name)
                if (gt) { const int t = axis[q]; axis[q] = axis[q + 1]; axis[q + 1] = t; }
            }
        int idx[3];                                   // idx[name] = position of that name in `axis`
        for (int p = 0; p < 3; ++p) idx[axis[p]] = p;
        double ctr[3];
        for (int r = 0; r < 3; ++r) {
            ctr[r] = s_mean[r];
            for (int k = 0; k < 3; ++k) ctr[r] += s_R[r][k] * 0.5 * ((double)elo[k] + (double)ehi[k]);
        }
        const float ext[3] = {ehi[0] - elo[0], ehi[1] - elo[1], ehi[2] - elo[2]};
        // R' = [R[:,idx z], R[:,idx y], R[:,idx x]]
        double Rp[3][3];
        for (int r = 0; r < 3; ++r) { Rp[r][0] = s_R[r][idx[2]]; Rp[r][1] = s_R[r][idx[1]]; Rp[r][2] = s_R[r][idx[0]]; }
        // scipy 1.11.4 Rotation.from_matrix (no determinant check; R' is left-handed for odd shuffles):
        // quaternion from the largest of (m00, m11, m22, trace), normalised; yaw = first 'zyx' angle
        // of that quaternion's rotation = atan2(2(zw - xy), 1 - 2(y^2 + z^2)).
        {
            const double tr = Rp[0][0] + Rp[1][1] + Rp[2][2];
            const double dec[4] = {Rp[0][0], Rp[1][1], Rp[2][2], tr};
            int choice = 0;
            for (int k = 1; k < 4; ++k)
                if (dec[k] > dec[choice]) choice = k;
            double q[4];
            if (choice != 3) {
                const int a = choice, b = (a + 1) % 3, cc = (b + 1) % 3;
                q[a] = 1.0 - tr + 2.0 * Rp[a][a];
                q[b] = Rp[b][a] + Rp[a][b];
                q[cc] = Rp[cc][a] + Rp[a][cc];
                q[3] = Rp[cc][b] - Rp[b][cc];
            } else {
                q[0] = Rp[2][1] - Rp[1][2];
                q[1] = Rp[0][2] - Rp[2][0];
                q[2] = Rp[1][0] - Rp[0][1];
                q[3] = 1.0 + tr;
            }
            const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            for (int k = 0; k < 4; ++k) q[k] /= nq;
            out[0] = (float)atan2(2.0 * (q[2] * q[3] - q[0] * q[1]), 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]));
        }
        out[1] = (float)ctr[0]; out[2] = (float)ctr[1]; out[3] = (float)ctr[2];
        out[4] = ext[idx[0]]; out[5] = ext[idx[1]]; out[6] = ext[idx[2]];
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) out[7 + 3 * r + k] = (float)Rp[r][k];
    }
}

}  // namespace cm3d

using namespace cm3d;

extern "C" int64_t cm3d_hull_obb_ws_words(int64_t seg_cap) { return 6 * seg_cap + (seg_cap + 3) / 4 + 4; }

extern "C" int cm3d_hull_obb(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                             int min_pts, int mode, const int32_t *order, int32_t *hull_ws, int64_t hull_ws_words,
                             float *obb, int32_t *hull_info, const int32_t *errflags, void *stream)
{
    if (n_inst_total < 0 || seg_cap < 0 || (mode != 0 && mode != 1)) return CM3D_EINVAL;
    if (n_inst_total == 0) return CM3D_OK;
    if (!seg_xyzw || !seg_off || !obb || !errflags) return CM3D_EINVAL;
    if (mode == 0 && (!hull_ws || hull_ws_words < cm3d_hull_obb_ws_words(seg_cap))) return CM3D_EINVAL;
    const size_t smem = mode == 0 ? sizeof(HullShared) : 0;
    if (smem) {
        cudaError_t e = cudaFuncSetAttribute(k_hull_obb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(1000 + (int)e);
    }
    k_hull_obb<<<n_inst_total, kHullThreads, smem, (cudaStream_t)stream>>>(seg_xyzw, seg_cap, seg_off, min_pts, mode, order,
                                                                           hull_ws, obb, hull_info, errflags);
    CM3D_LAUNCH_CHECK();
    return CM3D_OK;
}
