// GLSL-subset prelude for the generic hook runner (mpv_prescalers_b200/hookrunner.py): the vector types, built-ins and
// sampler functions an mpv user shader body expects, as CUDA device code.  Compiled by NVRTC together with the
// transpiled pass bodies (-default-device: everything here is a device function).  Arithmetic is plain IEEE float32 with
// FMA contraction off (--fmad=false), i.e. every GLSL operator is one rounded operation, which is what the CPU oracle
// (oracle/glsl_exec.py) does too.

typedef long long i64;

struct ivec2;
struct vec2 {
  float x, y;
  vec2() : x(0.f), y(0.f) {}
  explicit vec2(float s) : x(s), y(s) {}
  explicit vec2(const ivec2& v);
  vec2(float a, float b) : x(a), y(b) {}
  float& operator[](int i) { return i == 0 ? x : y; }
  float operator[](int i) const { return i == 0 ? x : y; }
};
typedef unsigned int uint;
struct uvec3 {   // the gl_* built-ins of a compute pass
  uint x, y, z;
  uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
struct ivec2 {
  int x, y;
  ivec2() : x(0), y(0) {}
  explicit ivec2(int s) : x(s), y(s) {}
  ivec2(int a, int b) : x(a), y(b) {}
  explicit ivec2(uvec3 v) : x((int)v.x), y((int)v.y) {}
  explicit ivec2(vec2 v) : x((int)v.x), y((int)v.y) {}
};
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(ivec2 a, ivec2 b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator*(ivec2 a, ivec2 b) { return ivec2(a.x * b.x, a.y * b.y); }
inline ivec2 operator+(ivec2 a, int b) { return ivec2(a.x + b, a.y + b); }
inline ivec2 operator-(ivec2 a, int b) { return ivec2(a.x - b, a.y - b); }
inline ivec2 operator*(ivec2 a, int b) { return ivec2(a.x * b, a.y * b); }
struct vec3 {
  float x, y, z;
  vec3() : x(0.f), y(0.f), z(0.f) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(vec2 a, float c) : x(a.x), y(a.y), z(c) {}
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
  float x, y, z, w;
  vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(vec3 a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
  vec4(vec2 a, vec2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
  vec4(vec2 a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  // the swizzles the hook files use (the transpiler turns `.wzyx` into `.wzyx()`)
  vec3 xyz() const { return vec3(x, y, z); }
  vec4 wzyx() const { return vec4(w, z, y, x); }
  vec2 xy() const { return vec2(x, y); }
  vec2 zw() const { return vec2(z, w); }
  vec2 wz() const { return vec2(w, z); }
  vec2 wx() const { return vec2(w, x); }
  vec2 zy() const { return vec2(z, y); }
  vec2 yx() const { return vec2(y, x); }
  vec2 xw() const { return vec2(x, w); }
  vec2 yz() const { return vec2(y, z); }
};

#define GLSL_VEC_OPS(V, ...)                                                                            \
  inline V operator+(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] + b[i]; return r; }   \
  inline V operator-(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] - b[i]; return r; }   \
  inline V operator*(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] * b[i]; return r; }   \
  inline V operator/(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] / b[i]; return r; }   \
  inline V operator+(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] + b; return r; }   \
  inline V operator-(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] - b; return r; }   \
  inline V operator*(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] * b; return r; }   \
  inline V operator/(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] / b; return r; }   \
  inline V operator+(float a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a + b[i]; return r; }   \
  inline V operator-(float a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a - b[i]; return r; }   \
  inline V operator*(float a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a * b[i]; return r; }   \
  inline V operator/(float a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a / b[i]; return r; }   \
  inline V operator-(V a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = -a[i]; return r; }               \
  inline V& operator+=(V& a, V b) { a = a + b; return a; }                                                  \
  inline V& operator-=(V& a, V b) { a = a - b; return a; }                                                  \
  inline V& operator*=(V& a, V b) { a = a * b; return a; }                                                  \
  inline V& operator/=(V& a, V b) { a = a / b; return a; }                                                  \
  inline V& operator+=(V& a, float b) { a = a + b; return a; }                                              \
  inline V& operator-=(V& a, float b) { a = a - b; return a; }                                              \
  inline V& operator*=(V& a, float b) { a = a * b; return a; }                                              \
  inline V& operator/=(V& a, float b) { a = a / b; return a; }                                              \
  inline V floor(V a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = floorf(a[i]); return r; }            \
  inline V fract(V a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] - floorf(a[i]); return r; }     \
  inline V abs(V a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = fabsf(a[i]); return r; }               \
  inline V sqrt(V a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = sqrtf(a[i]); return r; }              \
  inline V max(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = fmaxf(a[i], b[i]); return r; }    \
  inline V min(V a, V b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = fminf(a[i], b[i]); return r; }    \
  inline V max(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = fmaxf(a[i], b); return r; }   \
  inline V min(V a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = fminf(a[i], b); return r; }   \
  inline V clamp(V a, float lo, float hi) { return min(max(a, lo), hi); }                                   \
  inline V clamp(V a, V lo, V hi) { return min(max(a, lo), hi); }                                           \
  inline V mix(V a, V b, float t) { return a * (1.0f - t) + b * t; }                                        \
  inline V mix(V a, V b, V t) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r[i] = a[i] * (1.0f - t[i]) + b[i] * t[i]; return r; } \
  inline V mix(V a, V b, bool t) { return t ? b : a; }                                                     \
  inline float dot(V a, V b) { float s = a[0] * b[0]; for (int i = 1; i < __VA_ARGS__; ++i) s = s + a[i] * b[i]; return s; }

GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)

// compute-pass synchronisation
inline void barrier() { __syncthreads(); }
inline void groupMemoryBarrier() { __threadfence_block(); }
inline void memoryBarrierShared() { __threadfence_block(); }

// scalar built-ins (GLSL names)
inline float fract(float a) { return a - floorf(a); }
inline float floor(float a) { return floorf(a); }
inline float abs(float a) { return fabsf(a); }
inline float sqrt(float a) { return sqrtf(a); }
inline float inversesqrt(float a) { return 1.0f / sqrtf(a); }
inline float exp(float a) { return expf(a); }
inline float log2(float a) { return log2f(a); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline float clamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float mix(float a, float b, bool t) { return t ? b : a; }
inline float mod(float a, float b) { return a - b * floorf(a / b); }
inline float atan(float y, float x) { return atan2f(y, x); }
inline float intBitsToFloat(int v) { return __int_as_float(v); }

struct mat4x3 {   // 4 columns of vec3
  vec3 c[4];
  mat4x3() {}
  explicit mat4x3(float s) { c[0] = vec3(s, 0.f, 0.f); c[1] = vec3(0.f, s, 0.f); c[2] = vec3(0.f, 0.f, s); c[3] = vec3(0.f); }   // GLSL: s on the diagonal
  mat4x3(vec3 a, vec3 b, vec3 d, vec3 e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
  vec3& operator[](int i) { return c[i]; }
  vec3 operator[](int i) const { return c[i]; }
};
inline mat4x3 matrixCompMult(mat4x3 a, mat4x3 b) { return mat4x3(a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]); }
inline mat4x3 outerProduct(vec3 cv, vec4 rv) { return mat4x3(cv * rv.x, cv * rv.y, cv * rv.z, cv * rv.w); }
inline vec3 operator*(mat4x3 m, vec4 v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline mat4x3 operator+(mat4x3 a, mat4x3 b) { return mat4x3(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]); }
inline mat4x3& operator+=(mat4x3& a, mat4x3 b) { a = a + b; return a; }

// ---- textures ------------------------------------------------------------------------------------------------------
// A texture = `comps` interleaved float32 components per texel, `n` frames (stride 0 for LUTs).  Missing components read as
// (0, 0, 0, 1) like a GL sampler.  Addressing is clamp-to-edge.
struct Tex {
  const float* p;
  int w, h, comps, linear;
  i64 stride_n;
};
inline vec4 tex_fetch(const Tex& t, int frame, int ix, int iy) {
  ix = ix < 0 ? 0 : (ix > t.w - 1 ? t.w - 1 : ix);
  iy = iy < 0 ? 0 : (iy > t.h - 1 ? t.h - 1 : iy);
  const float* q = t.p + frame * t.stride_n + ((i64)iy * t.w + ix) * t.comps;
  vec4 r(0.f, 0.f, 0.f, 1.f);
  r.x = q[0];
  if (t.comps > 1) r.y = q[1];
  if (t.comps > 2) r.z = q[2];
  if (t.comps > 3) r.w = q[3];
  return r;
}
// texture(): NEAREST picks the texel that contains the coordinate; LINEAR blends the four texels around it exactly as
// the reference's sampler arithmetic does (u = c * size - 0.5)
inline vec4 tex_sample(const Tex& t, int frame, vec2 c) {
  if (!t.linear) return tex_fetch(t, frame, (int)floorf(c.x * (float)t.w), (int)floorf(c.y * (float)t.h));
  const float u = c.x * (float)t.w - 0.5f, v = c.y * (float)t.h - 0.5f;
  const float u0 = floorf(u), v0 = floorf(v);
  const float fu = u - u0, fv = v - v0;
  const int x0 = (int)u0, y0 = (int)v0;
  const vec4 top = tex_fetch(t, frame, x0, y0) * (1.0f - fu) + tex_fetch(t, frame, x0 + 1, y0) * fu;
  const vec4 bot = tex_fetch(t, frame, x0, y0 + 1) * (1.0f - fu) + tex_fetch(t, frame, x0 + 1, y0 + 1) * fu;
  return top * (1.0f - fv) + bot * fv;
}
// textureGatherOffset: the 2x2 footprint a LINEAR fetch at c would use, shifted by `off`; component `comp` of
// (x, y+1), (x+1, y+1), (x+1, y), (x, y) in .x .y .z .w.  GPUs resolve the footprint in fixed point (8 sub-texel bits), so
// a coordinate that sits on a texel centre up to float rounding selects that texel: emulated by the +1/512.
inline vec4 tex_gather(const Tex& t, int frame, vec2 c, ivec2 off, int comp) {
  const int x0 = (int)floorf(c.x * (float)t.w - 0.5f + 0.001953125f) + off.x;
  const int y0 = (int)floorf(c.y * (float)t.h - 0.5f + 0.001953125f) + off.y;
  return vec4(tex_fetch(t, frame, x0, y0 + 1)[comp], tex_fetch(t, frame, x0 + 1, y0 + 1)[comp],
              tex_fetch(t, frame, x0 + 1, y0)[comp], tex_fetch(t, frame, x0, y0)[comp]);
}
