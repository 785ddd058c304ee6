// Build scaffolding for oracle/_ref ONLY: no-op msgpack surface (msgpack-c is an un-vendored
// NuGet dependency of the reference, DeepLearning/packages.config:3; serialisation is off the path).
#pragma once
#include <cstddef>
namespace msgpack {
  struct object {};
  namespace type {
    template <class... A> struct define_array_stub {
      template <class P> void msgpack_pack(P&) const {}
      void msgpack_unpack(const object&) {}
    };
    template <class... A> define_array_stub<A...> make_define_array(A&...) { return {}; }
    template <class... A> define_array_stub<A...> make_define_array(const A&...) { return {}; }
  }
}
#define MSGPACK_DEFINE(...) template <class P> void msgpack_pack(P&) const {} void msgpack_unpack(const msgpack::object&) {}
#define MSGPACK_ADD_ENUM(x)
#define MSGPACK_BASE(b) (*const_cast<b*>(static_cast<const b*>(this)))
namespace msgpack {
  template <class S, class T> void pack(S&, const T&) {}
  struct object_handle_stub { struct o { template <class T> T as() const { return T{}; } }; o get() const { return {}; } };
  using unpacked = object_handle_stub;
  inline void unpack(unpacked&, const char*, std::size_t) {}
}
