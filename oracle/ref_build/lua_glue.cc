// TEST INFRASTRUCTURE (oracle/): the handful of Lua<->C++ conversions the
// reference's C++ core expects its (generated) binding layer to provide.
//
// The reference passes weight / gradient / component dictionaries between
// C++ components as AprilUtils::LuaTable (packages/basics/util/c_src/
// lua_table.h:100-760), whose per-type conversions are declared with
// DECLARE_LUA_TABLE_BIND_SPECIALIZATION (lua_table.h:51-58) and defined by the
// luabind-generated bindings.  The generated code is not built here; these are
// plain re-implementations: an object is stored as a Lua full userdata
// {type tag, pointer} that holds one reference, released by __gc.
#include <cstring>

extern "C" {
#include "lauxlib.h"
#include "lua.h"
#include "lualib.h"
}

#include "MersenneTwister.h"
#include "ann_component.h"
#include "complex_number.h"
#include "gpu_mirrored_memory_block.h"
#include "lua_table.h"
#include "matrixFloat.h"
#include "sparse_matrixFloat.h"
#include "token_base.h"

// mmapped_data.cc reads the commit number the reference's build stamps in.
const char* __COMMIT_NUMBER__ = "0";

namespace {

struct Box {
  const void* tag;
  Referenced* ref;  // for the release in __gc
  void* ptr;                    // the T* that was pushed
};

template <typename T>
struct Tag {
  static const char id;
};
template <typename T>
const char Tag<T>::id = 0;

int box_gc(lua_State* L) {
  Box* b = static_cast<Box*>(lua_touserdata(L, 1));
  if (b && b->ref) {
    DecRef(b->ref);
    b->ref = 0;
  }
  return 0;
}

template <typename T>
void push_boxed(lua_State* L, T* value) {
  if (value == 0) {
    lua_pushnil(L);
    return;
  }
  Box* b = static_cast<Box*>(lua_newuserdata(L, sizeof(Box)));
  b->tag = &Tag<T>::id;
  b->ref = value;
  b->ptr = value;
  IncRef(value);
  if (luaL_newmetatable(L, "b200.oracle.box")) {
    lua_pushcfunction(L, box_gc);
    lua_setfield(L, -2, "__gc");
  }
  lua_setmetatable(L, -2);
}

template <typename T>
bool is_boxed(lua_State* L, int idx) {
  if (lua_type(L, idx) != LUA_TUSERDATA) return false;
  Box* b = static_cast<Box*>(lua_touserdata(L, idx));
  return b && b->tag == &Tag<T>::id;
}

template <typename T>
T* to_boxed(lua_State* L, int idx) {
  return is_boxed<T>(L, idx) ? static_cast<T*>(static_cast<Box*>(lua_touserdata(L, idx))->ptr) : 0;
}

}  // namespace

#define B200_BOXED_TYPE(T)                                                        \
  namespace AprilUtils {                                                          \
  template <>                                                                     \
  T* LuaTable::convertTo<T*>(lua_State * L, int idx) {                            \
    return to_boxed<T>(L, idx);                                                   \
  }                                                                               \
  template <>                                                                     \
  void LuaTable::pushInto<T*>(lua_State * L, T * value) {                         \
    push_boxed<T>(L, value);                                                      \
  }                                                                               \
  template <>                                                                     \
  void LuaTable::pushInto<T>(lua_State * L, SharedPtr<T> value) {                 \
    push_boxed<T>(L, value.get());                                                \
  }                                                                               \
  template <>                                                                     \
  bool LuaTable::checkType<T*>(lua_State * L, int idx) {                          \
    return is_boxed<T>(L, idx);                                                   \
  }                                                                               \
  }

B200_BOXED_TYPE(Basics::MatrixFloat)
B200_BOXED_TYPE(Basics::SparseMatrixFloat)
B200_BOXED_TYPE(Basics::MTRand)
B200_BOXED_TYPE(ANN::ANNComponent)
B200_BOXED_TYPE(AprilMath::CharGPUMirroredMemoryBlock)
B200_BOXED_TYPE(AprilMath::FloatGPUMirroredMemoryBlock)
B200_BOXED_TYPE(AprilMath::DoubleGPUMirroredMemoryBlock)
B200_BOXED_TYPE(AprilMath::Int32GPUMirroredMemoryBlock)
B200_BOXED_TYPE(AprilMath::ComplexFGPUMirroredMemoryBlock)
B200_BOXED_TYPE(AprilMath::BoolGPUMirroredMemoryBlock)

namespace AprilUtils {

// A nested table travels by value: the registry reference is re-pushed.
template <>
LuaTable LuaTable::convertTo<LuaTable>(lua_State* L, int idx) {
  return LuaTable(L, idx);
}
template <>
void LuaTable::pushInto<LuaTable>(lua_State* L, LuaTable value) {
  value.pushTable(L);
}
template <>
bool LuaTable::checkType<LuaTable>(lua_State* L, int idx) {
  return lua_istable(L, idx);
}

typedef SharedPtr<Basics::Token> TokenPtr;
template <>
TokenPtr LuaTable::convertTo<TokenPtr>(lua_State* L, int idx) {
  return TokenPtr(to_boxed<Basics::Token>(L, idx));
}
template <>
void LuaTable::pushInto<Basics::Token>(lua_State* L, TokenPtr value) {
  push_boxed<Basics::Token>(L, value.get());
}
template <>
bool LuaTable::checkType<TokenPtr>(lua_State* L, int idx) {
  return is_boxed<Basics::Token>(L, idx);
}

// Complex numbers appear in serialisation helpers only: a {re, im} array.
template <>
AprilMath::ComplexF LuaTable::convertTo<AprilMath::ComplexF>(lua_State* L, int idx) {
  lua_rawgeti(L, idx, 1);
  lua_rawgeti(L, idx < 0 ? idx - 1 : idx, 2);
  AprilMath::ComplexF c((float)lua_tonumber(L, -2), (float)lua_tonumber(L, -1));
  lua_pop(L, 2);
  return c;
}
template <>
bool LuaTable::checkType<AprilMath::ComplexF>(lua_State* L, int idx) {
  return lua_istable(L, idx);
}

}  // namespace AprilUtils
