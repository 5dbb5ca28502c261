// vm.h — the vector maths layer of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
//
// Two interchangeable back ends behind the names the oracle uses (the glm/GLSL
// names vec3, dot, normalize, mix, ...):
//
//   default            : the small implementation below.  It restates, operation
//                        for operation, the arithmetic order of glm 0.9.8.5 built
//                        with GLM_FORCE_PURE (the maths library the reference's
//                        C++ side uses, vendored at <reference>/glm):
//                          dot(vec3)   = (x+y)+z                 glm/detail/func_geometric.inl:54-61
//                          dot(vec4)   = (x+y)+(z+w)             glm/detail/func_geometric.inl:64-71
//                          length      = sqrt(dot(v,v))          glm/detail/func_geometric.inl:14-20
//                          normalize   = v * (1/sqrt(dot(v,v)))  glm/detail/func_geometric.inl:88-96, func_exponential.inl:128-133
//                          mat4*vec4   = (m0v0+m1v1)+(m2v2+m3v3) glm/detail/type_mat4x4.inl:526-537
//                          mix(x,y,a)  = x + a*(y-x)             glm/detail/func_common.inl (compute_mix_*)
//                          clamp       = min(max(x,lo),hi)       glm/detail/func_common.inl:543-547
//                          min(x,y)    = x<y?x:y ; max = x>y?x:y glm/detail/func_common.inl:15-28
//   -DORACLE_USE_GLM   : the real glm headers (-I<reference root>), used by the
//                        oracle/_ref build to cross-check the restatement above
//                        bit for bit (tests/test_oracle_selfchecks.py).
//
// Both must be compiled with -ffp-contract=off so no FMA is formed.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef ORACLE_USE_GLM
#define GLM_FORCE_PURE
#include <glm/glm.hpp>
namespace vm = glm;
#else
namespace vm {

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
};
struct vec4;
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit vec3(const vec4& v);
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3& operator+=(const vec3& b) { x += b.x; y += b.y; z += b.z; return *this; }
    vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(const vec2& v, float z_, float w_) : x(v.x), y(v.y), z(z_), w(w_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec4& operator+=(const vec4& b) { x += b.x; y += b.y; z += b.z; w += b.w; return *this; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

struct mat4 {
    vec4 c[4];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator/(float s, vec2 a) { return vec2(s / a.x, s / a.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator+(float s, vec3 a) { return vec3(s + a.x, s + a.y, s + a.z); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }

inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, vec4 a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }

inline vec4 operator*(const mat4& m, const vec4& v) {
    const vec4 Mul0 = m[0] * vec4(v.x), Mul1 = m[1] * vec4(v.y);
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mul2 = m[2] * vec4(v.z), Mul3 = m[3] * vec4(v.w);
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

inline float min(float x, float y) { return x < y ? x : y; }
inline float max(float x, float y) { return x > y ? x : y; }
inline int min(int x, int y) { return x < y ? x : y; }
inline int max(int x, int y) { return x > y ? x : y; }
inline vec3 min(vec3 a, float s) { return vec3(min(a.x, s), min(a.y, s), min(a.z, s)); }
inline vec3 max(vec3 a, float s) { return vec3(max(a.x, s), max(a.y, s), max(a.z, s)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 clamp(vec3 a, float lo, float hi) { return vec3(clamp(a.x, lo, hi), clamp(a.y, lo, hi), clamp(a.z, lo, hi)); }
inline float abs(float x) { return std::fabs(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float exp(float x) { return std::exp(x); }
inline float pow(float x, float y) { return std::pow(x, y); }
inline float floor(float x) { return std::floor(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

inline float dot(vec3 a, vec3 b) { const vec3 t = a * b; return t.x + t.y + t.z; }
inline float dot(vec4 a, vec4 b) { const vec4 t = a * b; return (t.x + t.y) + (t.z + t.w); }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
inline float mix(float x, float y, float a) { return x + a * (y - x); }
inline vec3 mix(vec3 x, vec3 y, float a) { return x + a * (y - x); }

}  // namespace vm
#endif
