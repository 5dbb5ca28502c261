// scene_gen.cpp — procedural inputs for the VPL GI pass (host only, OpenMP).
//
// The reference produces the GI pass's inputs with its rasteriser (G-buffer pass,
// CSM depth pass, VPL cube shadow/diffuse pass — all out of scope, SURVEY.md §2.1
// rows 6-7).  This generator replaces them with analytic ray casts against
// axis-aligned boxes so that oracle, CUDA path and bench all see the same bytes:
//
//   G-buffer planes in the reference's formats (StratusRendererBackend.cpp:384-424)
//   VPL table (GpuVplData, StratusGpuCommon.h:217-232; radius = farPlane = 500,
//              StratusLight.h:320-323,354-356; intensity 1/50, CornellBox.cpp:182)
//   cube depth = saturate(dist/farPlane) and diffuse = albedo (shadowVpl.fs:25-33),
//              addressed through the inverse of GL's cube-face table
//   sun CSM: 4 orthographic cascades, split per StratusRendererFrontend.cpp:613-640,
//              cascade planes per :789-802.
//
// Scenes: kind 0 = Cornell box, kind 1 = Sponza-scale hall.
// Deterministic: PCG32 seeded by SgsDesc::seed.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>
#include <algorithm>
#include "../../include/stratusgi.h"

namespace {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double len(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 norm(V3 a) { double l = len(a); return a * (1.0 / l); }

struct Box {
    V3 lo, hi;
    uint8_t albedo[3];
    float roughness;
    bool insideOut;   // hit from the inside (enclosing shell): use the exit point, normals point inward
};

struct Hit { double t; int box; int axis; int sign; };   // normal = sign * e_axis

struct PCG32 {
    uint64_t state, inc;
    explicit PCG32(uint64_t seed, uint64_t seq = 0x5bd1e995u) { state = 0; inc = (seq << 1) | 1; next(); state += seed; next(); }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
    }
    double uniform() { return (next() >> 8) * (1.0 / 16777216.0); }
};

// IEEE binary32 -> binary16, round to nearest even (what GL does on a 16F store)
inline uint16_t f2h(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t em = x & 0x7fffffffu;
    if (em >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((em > 0x7f800000u) ? 0x200u : 0));
    if (em >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            // rounds to inf
    if (em < 0x33000001u) return (uint16_t)sign;                         // rounds to zero
    int e = (int)(em >> 23) - 127;
    uint32_t m = (em & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                   // subnormal half
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1))) q++;
    return (uint16_t)(sign | (base + q));
}

struct Mat4 { double m[16]; };   // column-major
inline Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; ++c) for (int rr = 0; rr < 4; ++rr) {
        double s = 0; for (int k = 0; k < 4; ++k) s += a.m[k * 4 + rr] * b.m[c * 4 + k];
        r.m[c * 4 + rr] = s;
    }
    return r;
}
bool invert(const Mat4& a, Mat4& out) {
    double inv[16]; const double* m = a.m;
    inv[0] = m[5]*m[10]*m[15]-m[5]*m[11]*m[14]-m[9]*m[6]*m[15]+m[9]*m[7]*m[14]+m[13]*m[6]*m[11]-m[13]*m[7]*m[10];
    inv[4] = -m[4]*m[10]*m[15]+m[4]*m[11]*m[14]+m[8]*m[6]*m[15]-m[8]*m[7]*m[14]-m[12]*m[6]*m[11]+m[12]*m[7]*m[10];
    inv[8] = m[4]*m[9]*m[15]-m[4]*m[11]*m[13]-m[8]*m[5]*m[15]+m[8]*m[7]*m[13]+m[12]*m[5]*m[11]-m[12]*m[7]*m[9];
    inv[12] = -m[4]*m[9]*m[14]+m[4]*m[10]*m[13]+m[8]*m[5]*m[14]-m[8]*m[6]*m[13]-m[12]*m[5]*m[10]+m[12]*m[6]*m[9];
    inv[1] = -m[1]*m[10]*m[15]+m[1]*m[11]*m[14]+m[9]*m[2]*m[15]-m[9]*m[3]*m[14]-m[13]*m[2]*m[11]+m[13]*m[3]*m[10];
    inv[5] = m[0]*m[10]*m[15]-m[0]*m[11]*m[14]-m[8]*m[2]*m[15]+m[8]*m[3]*m[14]+m[12]*m[2]*m[11]-m[12]*m[3]*m[10];
    inv[9] = -m[0]*m[9]*m[15]+m[0]*m[11]*m[13]+m[8]*m[1]*m[15]-m[8]*m[3]*m[13]-m[12]*m[1]*m[11]+m[12]*m[3]*m[9];
    inv[13] = m[0]*m[9]*m[14]-m[0]*m[10]*m[13]-m[8]*m[1]*m[14]+m[8]*m[2]*m[13]+m[12]*m[1]*m[10]-m[12]*m[2]*m[9];
    inv[2] = m[1]*m[6]*m[15]-m[1]*m[7]*m[14]-m[5]*m[2]*m[15]+m[5]*m[3]*m[14]+m[13]*m[2]*m[7]-m[13]*m[3]*m[6];
    inv[6] = -m[0]*m[6]*m[15]+m[0]*m[7]*m[14]+m[4]*m[2]*m[15]-m[4]*m[3]*m[14]-m[12]*m[2]*m[7]+m[12]*m[3]*m[6];
    inv[10] = m[0]*m[5]*m[15]-m[0]*m[7]*m[13]-m[4]*m[1]*m[15]+m[4]*m[3]*m[13]+m[12]*m[1]*m[7]-m[12]*m[3]*m[5];
    inv[14] = -m[0]*m[5]*m[14]+m[0]*m[6]*m[13]+m[4]*m[1]*m[14]-m[4]*m[2]*m[13]-m[12]*m[1]*m[6]+m[12]*m[2]*m[5];
    inv[3] = -m[1]*m[6]*m[11]+m[1]*m[7]*m[10]+m[5]*m[2]*m[11]-m[5]*m[3]*m[10]-m[9]*m[2]*m[7]+m[9]*m[3]*m[6];
    inv[7] = m[0]*m[6]*m[11]-m[0]*m[7]*m[10]-m[4]*m[2]*m[11]+m[4]*m[3]*m[10]+m[8]*m[2]*m[7]-m[8]*m[3]*m[6];
    inv[11] = -m[0]*m[5]*m[11]+m[0]*m[7]*m[9]+m[4]*m[1]*m[11]-m[4]*m[3]*m[9]-m[8]*m[1]*m[7]+m[8]*m[3]*m[5];
    inv[15] = m[0]*m[5]*m[10]-m[0]*m[6]*m[9]-m[4]*m[1]*m[10]+m[4]*m[2]*m[9]+m[8]*m[1]*m[6]-m[8]*m[2]*m[5];
    double det = m[0]*inv[0]+m[1]*inv[4]+m[2]*inv[8]+m[3]*inv[12];
    if (det == 0) return false;
    det = 1.0 / det;
    for (int i = 0; i < 16; ++i) out.m[i] = inv[i] * det;
    return true;
}
inline void xform(const Mat4& M, V3 p, double out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = M.m[r] * p.x + M.m[4 + r] * p.y + M.m[8 + r] * p.z + M.m[12 + r];
}

struct Camera { V3 pos, fwd, right, up; double tanHalfY, aspect; Mat4 pv, invPv; };

struct Scene {
    int kind, W, H, N, csmRes, cubeRes;
    uint32_t seed;
    double yawPerFrame;
    std::vector<Box> boxes;
    V3 camPos; double camYaw0;
    V3 sunDir; float sunColor[3];
    float minGiOcclusion;
    V3 sceneLo, sceneHi;        // interior bounds used for VPL seeding
    std::vector<SgiVplData> vpls;
    std::vector<SgiAtlasEntry> atlas;
};

const double kNear = 1.0, kFar = 1000.0, kFovY = 75.0;   // StratusRendererFrontend.h:27-28, StratusEngine.cpp:213

Hit trace(const Scene& s, V3 o, V3 d, double tmin = 1e-6) {
    Hit best{1e300, -1, 0, 0};
    const double od[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (size_t b = 0; b < s.boxes.size(); ++b) {
        const Box& bx = s.boxes[b];
        const double lo[3] = {bx.lo.x, bx.lo.y, bx.lo.z}, hi[3] = {bx.hi.x, bx.hi.y, bx.hi.z};
        double t0 = -1e300, t1 = 1e300; int a0 = 0, a1 = 0, s0 = 0, s1 = 0; bool miss = false;
        for (int a = 0; a < 3; ++a) {
            if (dd[a] == 0.0) { if (od[a] < lo[a] || od[a] > hi[a]) { miss = true; break; } continue; }
            double inv = 1.0 / dd[a];
            double ta = (lo[a] - od[a]) * inv, tb = (hi[a] - od[a]) * inv;
            int sa = -1, sb = 1;               // outward normal sign of the plane hit at ta / tb
            if (ta > tb) { std::swap(ta, tb); std::swap(sa, sb); }
            if (ta > t0) { t0 = ta; a0 = a; s0 = sa; }
            if (tb < t1) { t1 = tb; a1 = a; s1 = sb; }
        }
        if (miss || t0 > t1) continue;
        if (!bx.insideOut) {
            if (t0 > tmin && t0 < best.t) best = {t0, (int)b, a0, s0};
        } else {
            if (t1 > tmin && t1 < best.t) best = {t1, (int)b, a1, -s1};   // inward normal
        }
    }
    return best;
}

Camera makeCamera(const Scene& s, int frame) {
    Camera c;
    double yaw = s.camYaw0 + s.yawPerFrame * frame * (M_PI / 180.0);
    c.pos = s.camPos;
    c.fwd = {std::sin(yaw), 0.0, std::cos(yaw)};
    V3 worldUp{0, 1, 0};
    c.right = norm(cross(c.fwd, worldUp));
    c.up = cross(c.right, c.fwd);
    c.tanHalfY = std::tan(kFovY * 0.5 * M_PI / 180.0);
    c.aspect = double(s.W) / double(s.H);
    // view matrix (camera looks down -z in view space), column-major
    Mat4 V{}; V3 f = c.fwd, r = c.right, u = c.up;
    V.m[0] = r.x; V.m[4] = r.y; V.m[8] = r.z;  V.m[12] = -dot(r, c.pos);
    V.m[1] = u.x; V.m[5] = u.y; V.m[9] = u.z;  V.m[13] = -dot(u, c.pos);
    V.m[2] = -f.x; V.m[6] = -f.y; V.m[10] = -f.z; V.m[14] = dot(f, c.pos);
    V.m[3] = 0; V.m[7] = 0; V.m[11] = 0; V.m[15] = 1;
    Mat4 P{}; double g = 1.0 / c.tanHalfY;
    P.m[0] = g / c.aspect; P.m[5] = g; P.m[10] = -(kFar + kNear) / (kFar - kNear); P.m[11] = -1.0;
    P.m[14] = -2.0 * kFar * kNear / (kFar - kNear);
    c.pv = mul(P, V);
    invert(c.pv, c.invPv);
    return c;
}

void setAlbedo(Box& b, int r, int g, int bl) { b.albedo[0] = (uint8_t)r; b.albedo[1] = (uint8_t)g; b.albedo[2] = (uint8_t)bl; }

void addBox(Scene& s, V3 lo, V3 hi, int r, int g, int b, float rough, bool insideOut = false) {
    Box bx; bx.lo = lo; bx.hi = hi; setAlbedo(bx, r, g, b); bx.roughness = rough; bx.insideOut = insideOut;
    s.boxes.push_back(bx);
}

void buildCornell(Scene& s) {
    // room [-15,15]x[0,30]x[-15,15], open toward -z (CornellBox.cpp:83,186-294 for scale)
    const double th = 1.0;
    addBox(s, {-400, -60, -400}, {400, 400, 400}, 140, 150, 170, 0.9f, true);   // enclosing shell ("sky")
    addBox(s, {-200, -th, -200}, {200, 0, 200}, 160, 160, 160, 0.9f);           // ground / floor
    addBox(s, {-15 - th, 0, -15}, {-15, 30, 15}, 30, 200, 30, 0.7f);            // green wall x=-15
    addBox(s, {15, 0, -15}, {15 + th, 30, 15}, 200, 30, 30, 0.7f);              // red wall x=+15
    addBox(s, {-15 - th, 0, 15}, {15 + th, 30, 15 + th}, 200, 200, 200, 0.9f);  // back wall z=+15
    addBox(s, {-15 - th, 30, -15}, {15 + th, 30 + th, 15 + th}, 200, 200, 200, 0.9f); // ceiling
    addBox(s, {0, 0, -2}, {6, 18, 4}, 190, 190, 120, 0.5f);                     // tall box
    addBox(s, {-10, 0, -11}, {-3, 10, -4}, 120, 160, 210, 0.5f);                // short box
    s.camPos = {0, 15, -45}; s.camYaw0 = 0.0;
    s.sceneLo = {-15, 0, -15}; s.sceneHi = {15, 30, 15};
    s.minGiOcclusion = 0.65f;   // CornellBox.cpp:96
}

void buildHall(Scene& s) {
    // hall [-30,60]x[0,240]x[-140,180] (extent of the VPL grid, Sponza.cpp:176-178)
    const double th = 2.0;
    addBox(s, {-450, -100, -450}, {450, 600, 450}, 150, 170, 200, 0.9f, true);
    addBox(s, {-300, -th, -300}, {300, 0, 300}, 170, 160, 140, 0.7f);                      // floor
    addBox(s, {-30 - th, 0, -140}, {-30, 240, 180}, 190, 170, 150, 0.9f);                  // wall x=-30
    addBox(s, {60, 0, -140}, {60 + th, 240, 180}, 190, 170, 150, 0.9f);                    // wall x=60
    addBox(s, {-30 - th, 0, 180}, {60 + th, 240, 180 + th}, 180, 180, 170, 0.9f);          // end wall
    // two rows of 11 box columns
    for (int i = 0; i < 11; ++i) {
        double z = -125.0 + 28.0 * i;
        addBox(s, {-12, 0, z}, {-6, 90, z + 6}, 200, 190, 170, 0.5f);
        addBox(s, {36, 0, z}, {42, 90, z + 6}, 200, 190, 170, 0.5f);
    }
    // upper gallery slabs along both walls, and a partial roof leaving an atrium slot for the sun
    addBox(s, {-30, 90, -140}, {-6, 96, 180}, 150, 90, 70, 0.7f);
    addBox(s, {36, 90, -140}, {60, 96, 180}, 150, 90, 70, 0.7f);
    addBox(s, {-30 - th, 240, -140}, {0, 240 + th, 180}, 170, 170, 170, 0.9f);
    // banners (coloured boxes) for colour bleeding
    addBox(s, {-5, 40, -60}, {-4, 80, -40}, 200, 40, 40, 0.7f);
    addBox(s, {34, 40, 20}, {35, 80, 40}, 40, 80, 200, 0.7f);
    addBox(s, {-5, 40, 90}, {-4, 80, 110}, 40, 180, 60, 0.7f);
    s.camPos = {15, 30, -132}; s.camYaw0 = 0.0;
    s.sceneLo = {-30, 0, -140}; s.sceneHi = {60, 240, 180};
    s.minGiOcclusion = 0.95f;
}

void lightBasis(const Scene& s, V3& right, V3& up, V3& fwd) {
    fwd = s.sunDir;
    V3 ref = std::fabs(fwd.y) < 0.99 ? V3{0, 1, 0} : V3{1, 0, 0};
    right = norm(cross(ref, fwd));
    up = cross(fwd, right);
}

struct Cascade { double cx, cy, dk, zlo, zhi; };

void cascadeSplits(double ends[4], double begins[4]) {
    // StratusRendererFrontend.cpp:613-640
    const float znear = 1.0f, zfar = (float)kFar, lambda = 0.5f;
    const float clipRange = zfar - znear, ratio = zfar / znear;
    for (int i = 0; i < 4; ++i) {
        const float p = (i + 1) / 4.0f;
        const float lg = znear * std::pow(ratio, p);
        const float un = znear + clipRange * p;
        ends[i] = std::floor(lambda * lg + (1.0f - lambda) * un);
    }
    begins[0] = 0.0; for (int i = 1; i < 4; ++i) begins[i] = ends[i - 1] - 4.0;
}

void makeCascades(const Scene& s, const Camera& cam, Cascade out[4]) {
    double ends[4], begins[4]; cascadeSplits(ends, begins);
    V3 lr, lu, lf; lightBasis(s, lr, lu, lf);
    // light-space z range of the whole scene (all occluders)
    double zlo = 1e300, zhi = -1e300;
    for (size_t b = 1; b < s.boxes.size(); ++b) {   // skip the enclosing shell
        const Box& bx = s.boxes[b];
        for (int k = 0; k < 8; ++k) {
            V3 p{(k & 1) ? bx.hi.x : bx.lo.x, (k & 2) ? bx.hi.y : bx.lo.y, (k & 4) ? bx.hi.z : bx.lo.z};
            if (std::fabs(p.x) > 250 || std::fabs(p.z) > 250) continue;   // ignore the far ends of the ground slab
            double z = dot(p, lf); zlo = std::min(zlo, z); zhi = std::max(zhi, z);
        }
    }
    zlo -= 10.0; zhi += 10.0;
    for (int i = 0; i < 4; ++i) {
        const double ak = begins[i], bk = ends[i];
        const double xn = ak * cam.aspect * cam.tanHalfY, xf = bk * cam.aspect * cam.tanHalfY;
        const double yn = ak * cam.tanHalfY, yf = bk * cam.tanHalfY;
        V3 corners[8]; int n = 0;
        for (int far_ = 0; far_ < 2; ++far_) for (int sy = -1; sy <= 1; sy += 2) for (int sx = -1; sx <= 1; sx += 2) {
            double d = far_ ? bk : ak, x = (far_ ? xf : xn) * sx, y = (far_ ? yf : yn) * sy;
            corners[n++] = cam.pos + cam.fwd * d + cam.right * x + cam.up * y;
        }
        double maxLen = 0; for (int a = 0; a < 8; ++a) for (int b = a + 1; b < 8; ++b) maxLen = std::max(maxLen, len(corners[a] - corners[b]));
        const double dk = std::ceil(maxLen);
        double minX = 1e300, maxX = -1e300, minY = 1e300, maxY = -1e300;
        for (int a = 0; a < 8; ++a) {
            double x = dot(corners[a], lr), y = dot(corners[a], lu);
            minX = std::min(minX, x); maxX = std::max(maxX, x); minY = std::min(minY, y); maxY = std::max(maxY, y);
        }
        const double T = dk / s.csmRes;   // texel size; snap the centre for stability like FE:730-732
        out[i].cx = std::floor((maxX + minX) / (2.0 * T)) * T;
        out[i].cy = std::floor((maxY + minY) / (2.0 * T)) * T;
        out[i].dk = dk; out[i].zlo = zlo; out[i].zhi = zhi;
    }
}

void cascadeMatrix(const Scene& s, const Cascade& c, float out[16]) {
    V3 lr, lu, lf; lightBasis(s, lr, lu, lf);
    const double sx = 2.0 / c.dk, sz = 2.0 / (c.zhi - c.zlo);
    double m[16] = {
        lr.x * sx, lu.x * sx, lf.x * sz, 0,
        lr.y * sx, lu.y * sx, lf.y * sz, 0,
        lr.z * sx, lu.z * sx, lf.z * sz, 0,
        -c.cx * sx, -c.cy * sx, -c.zlo * sz - 1.0, 1};
    for (int i = 0; i < 16; ++i) out[i] = (float)m[i];
}

void generateVpls(Scene& s) {
    PCG32 rng(s.seed, 77);
    s.vpls.assign(s.N, SgiVplData{});
    s.atlas.assign(s.N, SgiAtlasEntry{});
    V3 lr, lu, lf; lightBasis(s, lr, lu, lf);
    const V3 lo = s.sceneLo, hi = s.sceneHi;
    const V3 ctr = (lo + hi) * 0.5; const double rad = len(hi - lo) * 0.5;
    int made = 0, guard = 0;
    while (made < s.N && guard < s.N * 1000) {
        ++guard;
        Hit h; V3 o, d;
        if (rng.uniform() < 0.7) {
            // first hit seen from the sun: lit by construction
            double u = (rng.uniform() * 2 - 1) * rad, v = (rng.uniform() * 2 - 1) * rad;
            o = ctr + lr * u + lu * v - lf * (rad + 300.0); d = lf;
        } else {
            o = {lo.x + (hi.x - lo.x) * rng.uniform(), lo.y + (hi.y - lo.y) * rng.uniform(), lo.z + (hi.z - lo.z) * rng.uniform()};
            double z = rng.uniform() * 2 - 1, ph = rng.uniform() * 2 * M_PI, r = std::sqrt(std::max(0.0, 1 - z * z));
            d = {r * std::cos(ph), z, r * std::sin(ph)};
        }
        h = trace(s, o, d);
        if (h.box <= 0) continue;                       // missed or hit the enclosing shell
        V3 p = o + d * h.t;
        if (p.x < lo.x - 1 || p.x > hi.x + 1 || p.y < lo.y - 1 || p.y > hi.y + 1 || p.z < lo.z - 1 || p.z > hi.z + 1) continue;
        double nn[3] = {0, 0, 0}; nn[h.axis] = h.sign;
        p = p + V3{nn[0], nn[1], nn[2]} * 1.0;         // 1 unit off the surface
        SgiVplData& v = s.vpls[made];
        v.position[0] = (float)p.x; v.position[1] = (float)p.y; v.position[2] = (float)p.z; v.position[3] = 1.0f;
        v.radius = 500.0f; v.farPlane = 500.0f; v.intensity = 1.0f / 50.0f;
        ++made;
    }
    // scattered atlas assignment (an LRU cache hands out arbitrary free slots, BE:2606-2636):
    // a fixed permutation of the first max(N, 320)-rounded slots
    const int slots = std::min(SGI_MAX_TOTAL_SHADOW_MAPS, ((s.N + SGI_MAX_TOTAL_SHADOWS_PER_ATLAS - 1) / SGI_MAX_TOTAL_SHADOWS_PER_ATLAS) * SGI_MAX_TOTAL_SHADOWS_PER_ATLAS);
    std::vector<int> perm(slots); for (int i = 0; i < slots; ++i) perm[i] = i;
    for (int i = slots - 1; i > 0; --i) { int j = (int)(rng.next() % (uint32_t)(i + 1)); std::swap(perm[i], perm[j]); }
    for (int i = 0; i < s.N; ++i) {
        int slot = perm[i % slots];
        s.atlas[i].index = slot / SGI_MAX_TOTAL_SHADOWS_PER_ATLAS;
        s.atlas[i].layer = slot % SGI_MAX_TOTAL_SHADOWS_PER_ATLAS;
    }
}

inline V3 cubeDir(int face, double sc, double tc) {
    switch (face) {       // inverse of GL's cube-face selection table
        case 0: return {1, -tc, -sc};
        case 1: return {-1, -tc, sc};
        case 2: return {sc, 1, tc};
        case 3: return {sc, -1, -tc};
        case 4: return {sc, -tc, 1};
        default: return {-sc, -tc, -1};
    }
}

}  // namespace

extern "C" {

typedef struct SgsDesc {
    int32_t kind;        // 0 Cornell box, 1 Sponza-scale hall
    int32_t width, height;
    int32_t num_vpls;
    int32_t csm_res;
    int32_t cube_res;
    uint32_t seed;
    float yaw_deg_per_frame;
} SgsDesc;

void* sgs_create(const SgsDesc* d) {
    if (!d || d->width <= 0 || d->height <= 0 || d->num_vpls < 0 || d->num_vpls > SGI_MAX_TOTAL_VPLS_BEFORE_CULLING) return nullptr;
    Scene* s = new Scene();
    s->kind = d->kind; s->W = d->width; s->H = d->height; s->N = d->num_vpls;
    s->csmRes = d->csm_res; s->cubeRes = d->cube_res; s->seed = d->seed; s->yawPerFrame = d->yaw_deg_per_frame;
    s->sunDir = norm(V3{0.35, -0.85, 0.4});
    const float sc[3] = {251.0f / 255.0f, 144.0f / 255.0f, 98.0f / 255.0f};     // CornellBox.cpp:57-59
    for (int i = 0; i < 3; ++i) s->sunColor[i] = sc[i] * 9.5f;
    if (s->kind == 0) buildCornell(*s); else buildHall(*s);
    generateVpls(*s);
    return s;
}

void sgs_destroy(void* p) { delete (Scene*)p; }

int sgs_num_boxes(void* p) { return (int)((Scene*)p)->boxes.size(); }

int sgs_frame_params(void* p, int frame, SgiFrameParams* out) {
    Scene& s = *(Scene*)p;
    std::memset(out, 0, sizeof(*out));
    Camera cam = makeCamera(s, frame), prev = makeCamera(s, frame - 1);
    for (int i = 0; i < 16; ++i) { out->invProjectionView[i] = (float)cam.invPv.m[i]; out->prevInvProjectionView[i] = (float)prev.invPv.m[i]; }
    out->viewPosition[0] = (float)cam.pos.x; out->viewPosition[1] = (float)cam.pos.y; out->viewPosition[2] = (float)cam.pos.z;
    out->time = 1000.0f + 16.0f * frame;
    out->deltaSeconds = 1.0 / 60.0;
    out->minGiOcclusionFactor = s.minGiOcclusion;
    out->minRoughness = 0.08f; out->usePerceptualRoughness = 1;
    out->fogDensity = 0.0f; out->fogColor[0] = out->fogColor[1] = out->fogColor[2] = 0.5f;
    out->infiniteLightDirection[0] = (float)s.sunDir.x; out->infiniteLightDirection[1] = (float)s.sunDir.y; out->infiniteLightDirection[2] = (float)s.sunDir.z;
    for (int i = 0; i < 3; ++i) out->infiniteLightColor[i] = s.sunColor[i];
    Cascade cs[4]; makeCascades(s, cam, cs);
    for (int i = 0; i < 4; ++i) cascadeMatrix(s, cs[i], out->cascadeProjViews[i]);
    double ends[4], begins[4]; cascadeSplits(ends, begins);
    for (int i = 1; i < 4; ++i) {   // FE:789-802: fk = (n, dot(-n,c) - ak) / (b_{k-1} - ak)
        const double ak = begins[i], inv = 1.0 / (ends[i - 1] - ak);
        out->cascadePlanes[i - 1][0] = (float)(cam.fwd.x * inv); out->cascadePlanes[i - 1][1] = (float)(cam.fwd.y * inv);
        out->cascadePlanes[i - 1][2] = (float)(cam.fwd.z * inv); out->cascadePlanes[i - 1][3] = (float)((-dot(cam.fwd, cam.pos) - ak) * inv);
    }
    out->totalNumLights = s.N;
    out->gatherMode = SGI_GATHER_STOCHASTIC;
    return 0;
}

// G-buffer rows [y0, y0+rows) of frame `frame`; any output pointer may be NULL.
int sgs_gbuffer(void* p, int frame, int y0, int rows, float* depth, uint8_t* normal4, uint8_t* albedo4, uint8_t* refl2,
                uint8_t* rma4, uint16_t* ssaoH, uint16_t* velocityH2, float* id, uint16_t* screenH4) {
    Scene& s = *(Scene*)p;
    Camera cam = makeCamera(s, frame), prev = makeCamera(s, frame - 1);
    const int W = s.W, H = s.H;
    #pragma omp parallel for schedule(dynamic, 4)
    for (int ly = 0; ly < rows; ++ly) {
        const int y = y0 + ly;
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)ly * W + x;
            const double u = (x + 0.5) / W, v = (y + 0.5) / H;
            V3 d = norm(cam.fwd + cam.right * ((u * 2 - 1) * cam.tanHalfY * cam.aspect) + cam.up * ((v * 2 - 1) * cam.tanHalfY));
            Hit h = trace(s, cam.pos, d);
            V3 hp = cam.pos + d * h.t;
            const Box& bx = s.boxes[h.box];
            double clip[4]; xform(cam.pv, hp, clip);
            double z01 = (clip[2] / clip[3]) * 0.5 + 0.5; z01 = std::min(1.0, std::max(0.0, z01));
            const double q = std::floor(z01 * 16777215.0 + 0.5);               // 24-bit unorm depth
            if (depth) depth[i] = (float)(q / 16777215.0);
            double nn[3] = {0, 0, 0}; nn[h.axis] = h.sign;
            if (normal4) for (int c = 0; c < 4; ++c) normal4[i * 4 + c] = c < 3 ? (uint8_t)std::floor((nn[c] * 0.5 + 0.5) * 255.0 + 0.5) : 255;
            if (albedo4) { for (int c = 0; c < 3; ++c) albedo4[i * 4 + c] = bx.albedo[c]; albedo4[i * 4 + 3] = 0; }
            if (refl2) {   // pbr_geometry_pass.fs:129-132 with reflectance 0.5, metallic 0, maxReflectivity 0.8
                const double refl = 0.5 + (0.8 - 0.5) * ((1.0 - bx.roughness) * 0.5);
                refl2[i * 2] = (uint8_t)std::floor(refl * 255.0 + 0.5); refl2[i * 2 + 1] = 0;
            }
            if (rma4) { rma4[i * 4] = (uint8_t)std::floor(bx.roughness * 255.0 + 0.5); rma4[i * 4 + 1] = 0; rma4[i * 4 + 2] = 0; rma4[i * 4 + 3] = 0; }
            if (ssaoH) {   // smooth, deterministic occlusion in [0.25,1] so the clamp(., 0.35, 1) is exercised
                const double ao = 0.25 + 0.75 * (0.5 + 0.5 * std::sin(0.013 * x) * std::cos(0.017 * y));
                ssaoH[i] = f2h((float)ao);
            }
            if (velocityH2) {
                double pc[4]; xform(prev.pv, hp, pc);
                const double cu = (clip[0] / clip[3]) * 0.5 + 0.5, cv = (clip[1] / clip[3]) * 0.5 + 0.5;
                const double pu = (pc[0] / pc[3]) * 0.5 + 0.5, pvv = (pc[1] / pc[3]) * 0.5 + 0.5;
                velocityH2[i * 2] = f2h((float)(cu - pu)); velocityH2[i * 2 + 1] = f2h((float)(cv - pvv));
            }
            if (id) id[i] = (float)h.box;
            if (screenH4) {   // stand-in for the deferred direct-lighting result (lightingColorBuffer)
                V3 n{nn[0], nn[1], nn[2]};
                double ndl = std::max(0.0, -dot(n, s.sunDir));
                double vis = 0.0;
                if (ndl > 0) { Hit sh = trace(s, hp + n * 1e-3, s.sunDir * -1.0); vis = (sh.box <= 0) ? 1.0 : 0.0; }
                for (int c = 0; c < 3; ++c) screenH4[i * 4 + c] = f2h((float)((bx.albedo[c] / 255.0) * (1.0 / M_PI) * ndl * vis * s.sunColor[c]));
                screenH4[i * 4 + 3] = f2h(1.0f);
            }
        }
    }
    return 0;
}

// The scene's geometry as the box list sgi_render_vpl_cubemaps consumes (float bounds; every bound of the procedural scenes is
// exactly representable).  Returns the number of boxes; `out` may be NULL to query it.
int sgs_boxes(void* p, SgiSceneBox* out, int capacity) {
    Scene& s = *(Scene*)p;
    const int n = (int)s.boxes.size();
    if (out) for (int b = 0; b < n && b < capacity; ++b) {
        const Box& bx = s.boxes[b];
        out[b].lo[0] = (float)bx.lo.x; out[b].lo[1] = (float)bx.lo.y; out[b].lo[2] = (float)bx.lo.z;
        out[b].hi[0] = (float)bx.hi.x; out[b].hi[1] = (float)bx.hi.y; out[b].hi[2] = (float)bx.hi.z;
        for (int c = 0; c < 3; ++c) out[b].albedo[c] = bx.albedo[c];
        out[b].albedo[3] = 255;
        out[b].insideOut = bx.insideOut ? 1 : 0;
    }
    return n;
}

int sgs_vpls(void* p, SgiVplData* out, SgiAtlasEntry* atlas) {
    Scene& s = *(Scene*)p;
    if (out) std::memcpy(out, s.vpls.data(), sizeof(SgiVplData) * s.N);
    if (atlas) std::memcpy(atlas, s.atlas.data(), sizeof(SgiAtlasEntry) * s.N);
    return 0;
}

// Cube maps of VPLs [first, first+count) written at their atlas slots.  `depth`/`diffuse4`
// are the concatenation of all atlases: slot = atlas*320 + layer, each slot 6*res*res texels.
int sgs_cubemaps(void* p, int first, int count, uint16_t* depth, uint8_t* diffuse4) {
    Scene& s = *(Scene*)p;
    const int R = s.cubeRes;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int k = first; k < first + count; ++k) {
        const SgiVplData& v = s.vpls[k];
        const size_t slot = (size_t)s.atlas[k].index * SGI_MAX_TOTAL_SHADOWS_PER_ATLAS + s.atlas[k].layer;
        const V3 o{v.position[0], v.position[1], v.position[2]};
        for (int f = 0; f < 6; ++f) for (int j = 0; j < R; ++j) for (int i = 0; i < R; ++i) {
            const double sc = 2.0 * (i + 0.5) / R - 1.0, tc = 2.0 * (j + 0.5) / R - 1.0;
            V3 d = norm(cubeDir(f, sc, tc));
            Hit h = trace(s, o, d);
            const size_t t = ((slot * 6 + f) * R + j) * R + i;
            const double dist = std::min(1.0, std::max(0.0, h.t / (double)v.farPlane));
            if (depth) depth[t] = (uint16_t)std::floor(dist * 65535.0 + 0.5);
            if (diffuse4) { const Box& bx = s.boxes[h.box]; for (int c = 0; c < 3; ++c) diffuse4[t * 4 + c] = bx.albedo[c]; diffuse4[t * 4 + 3] = 255; }
        }
    }
    return 0;
}

int sgs_csm(void* p, int frame, float* depth) {
    Scene& s = *(Scene*)p;
    Camera cam = makeCamera(s, frame);
    Cascade cs[4]; makeCascades(s, cam, cs);
    V3 lr, lu, lf; lightBasis(s, lr, lu, lf);
    const int R = s.csmRes;
    for (int c = 0; c < 4; ++c) {
        #pragma omp parallel for schedule(dynamic, 8)
        for (int j = 0; j < R; ++j) for (int i = 0; i < R; ++i) {
            const double lx = cs[c].cx + ((i + 0.5) / R - 0.5) * cs[c].dk, ly = cs[c].cy + ((j + 0.5) / R - 0.5) * cs[c].dk;
            V3 o = lr * lx + lu * ly + lf * cs[c].zlo;
            Hit h = trace(s, o, lf);
            double z01 = 1.0;
            if (h.box > 0) z01 = std::min(1.0, h.t / (cs[c].zhi - cs[c].zlo));
            const double q = std::floor(z01 * 16777215.0 + 0.5);
            depth[((size_t)c * R + j) * R + i] = (float)(q / 16777215.0);
        }
    }
    return 0;
}

// Exact (ray-traced) sun visibility of every VPL position: an independent check of the CSM chain.
int sgs_sun_visibility(void* p, uint8_t* out) {
    Scene& s = *(Scene*)p;
    for (int k = 0; k < s.N; ++k) {
        const SgiVplData& v = s.vpls[k];
        Hit h = trace(s, V3{v.position[0], v.position[1], v.position[2]}, s.sunDir * -1.0);
        out[k] = (h.box <= 0) ? 1 : 0;
    }
    return 0;
}

uint16_t sgs_float_to_half(float f) { return f2h(f); }

// The engine-side constant input of the SSAO pass: InitSSAO_ (StratusRendererBackend.cpp:759-779) shuffles k = 0..15 with a
// default-constructed std::default_random_engine and stores (cos, sin, 0) of 2*pi*k/16 in a 4x4 RGB16F texture.  out: 16 x 3 floats.
void sgs_ssao_rotation_table(float* out) {
    std::vector<float> ks(16);
    std::iota(ks.begin(), ks.end(), 0.0f);
    std::shuffle(ks.begin(), ks.end(), std::default_random_engine{});
    for (size_t i = 0; i < ks.size(); ++i) {
        const float r = 2.0f * float(M_PI) * ks[i] / 16.0f;
        out[i * 3] = std::cos(r); out[i * 3 + 1] = std::sin(r); out[i * 3 + 2] = 0.0f;
    }
}
double sgs_fovy_degrees(void) { return kFovY; }

}  // extern "C"
