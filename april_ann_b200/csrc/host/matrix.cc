// Matrix (device tensor), MTRand and small helpers of the host mirror.
#include <math.h>
#include <string.h>

#include <regex>

#include "ann.h"

namespace b200 {

void check(int status) {
  if (status != B200_OK) throw Error(status, b200_last_error_string());
}

// ------------------------------------------------------------------ MTRand
// Published MT19937 (Matsumoto & Nishimura) with the reference's accessors:
// packages/basics/random/c_src/MersenneTwister.cc:104-134 (randInt), :62-68 (rand),
// :226-255 (initialize/reload), :279-288 (shuffle).
void MTRand::seed(uint32_t s) {
  state[0] = s;
  for (int i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + (uint32_t)i;
  reload();
}
void MTRand::reload() {
  auto twist = [](uint32_t m, uint32_t s0, uint32_t s1) {
    uint32_t mix = (s0 & 0x80000000u) | (s1 & 0x7fffffffu);
    return m ^ (mix >> 1) ^ ((s1 & 1u) ? 0x9908b0dfu : 0u);
  };
  int i;
  for (i = 0; i < 624 - 397; ++i) state[i] = twist(state[i + 397], state[i], state[i + 1]);
  for (; i < 623; ++i) state[i] = twist(state[i + 397 - 624], state[i], state[i + 1]);
  state[623] = twist(state[396], state[623], state[0]);
  left = 624;
  pos = 0;
}
uint32_t MTRand::randInt() {
  if (left == 0) reload();
  --left;
  uint32_t s1 = state[pos++];
  s1 ^= (s1 >> 11);
  s1 ^= (s1 << 7) & 0x9d2c5680u;
  s1 ^= (s1 << 15) & 0xefc60000u;
  return s1 ^ (s1 >> 18);
}
uint32_t MTRand::randInt(uint32_t n) {
  uint32_t used = n;
  used |= used >> 1; used |= used >> 2; used |= used >> 4; used |= used >> 8; used |= used >> 16;
  uint32_t i;
  do i = randInt() & used; while (i > n);
  return i;
}
double MTRand::rand(double n) { return double(randInt()) * (1.0 / 4294967295.0) * n; }
void MTRand::shuffle(int size, int *v) {
  for (int i = 0; i < size; ++i) v[i] = i;
  for (int i = size - 1; i > 0; --i) {
    int j = (int)randInt((uint32_t)i);
    int s = v[i]; v[i] = v[j]; v[j] = s;
  }
}

// ------------------------------------------------------------------ Matrix
size_t Matrix::size() const {
  size_t n = 1;
  for (int d : dims) n *= (size_t)d;
  return n;
}
MatrixPtr Matrix::create(b200_ctx *ctx, const std::vector<int> &dims) {
  MatrixPtr m(new Matrix());
  m->ctx = ctx;
  m->dims = dims;
  for (int d : dims)
    if (d <= 0) throw Error(B200_ERR_BAD_ARG, "Matrix::create: dimensions must be positive");
  void *p = nullptr;
  check(b200_malloc(ctx, &p, m->size() * sizeof(float)));
  m->data = (float *)p;
  m->owns = true;
  registerCapturedMatrix(m);
  return m;
}
MatrixPtr Matrix::view(const MatrixPtr &parent, size_t offset, const std::vector<int> &dims) {
  MatrixPtr m(new Matrix());
  m->ctx = parent->ctx;
  m->dims = dims;
  if (offset + m->size() > parent->size()) throw Error(B200_ERR_BAD_ARG, "Matrix::view: out of range");
  m->data = parent->data + offset;
  m->parent = parent;
  return m;
}
MatrixPtr Matrix::wrap(b200_ctx *ctx, float *ptr, const std::vector<int> &dims) {
  MatrixPtr m(new Matrix());
  m->ctx = ctx;
  m->dims = dims;
  m->data = ptr;
  return m;
}
Matrix::~Matrix() {
  if (owns && data) b200_free(ctx, data);
}
MatrixPtr Matrix::rewrap(const std::vector<int> &new_dims) {
  size_t n = 1;
  for (int d : new_dims) n *= (size_t)d;
  if (n != size()) throw Error(B200_ERR_BAD_ARG, "rewrap: incompatible sizes");
  MatrixPtr m(new Matrix());
  m->ctx = ctx;
  m->dims = new_dims;
  m->data = data;
  // keep the owning block alive
  m->parent = parent ? parent : shared_from_this();
  return m;
}
void Matrix::zeros() { check(b200_memset_zero(ctx, data, size() * sizeof(float))); }
void Matrix::fromHost(const float *src) { check(b200_memcpy_h2d(ctx, data, src, size() * sizeof(float))); }
void Matrix::toHost(float *dst) const {
  check(b200_memcpy_d2h(ctx, dst, data, size() * sizeof(float)));
  check(b200_sync(ctx));
}
void Matrix::copyFrom(const Matrix &o) {
  if (o.size() != size()) throw Error(B200_ERR_BAD_ARG, "copyFrom: size mismatch");
  check(b200_memcpy_d2d(ctx, data, o.data, size() * sizeof(float)));
}

// The handful of Lua patterns the reference scripts pass to set_layerwise_option /
// randomize_weights ("b.", ".*w.*", "w1", "^b%d+$") mean the same as ECMAScript regexes
// once '%' is read as the escape character.
bool luaPatternMatch(const std::string &pattern, const std::string &s) {
  std::string rx;
  for (size_t i = 0; i < pattern.size(); ++i) {
    char c = pattern[i];
    if (c == '%' && i + 1 < pattern.size()) {
      char n = pattern[++i];
      if (n == 'd') rx += "[0-9]";
      else if (n == 'a') rx += "[A-Za-z]";
      else if (n == 'w') rx += "[A-Za-z0-9]";
      else if (n == 's') rx += "\\s";
      else { rx += '\\'; rx += n; }
    } else if (c == '-') {
      rx += "*?";
    } else {
      rx += c;
    }
  }
  return std::regex_search(s, std::regex(rx));
}

}  // namespace b200
