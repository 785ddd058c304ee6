// Build scaffolding for oracle/_ref ONLY: serial stand-in for MSVC's PPL parallel_reduce (Net.inl:420,452).
#pragma once
namespace concurrency {
  template <class It, class T, class F, class R>
  T parallel_reduce(It b, It e, const T& init, const F& f, const R&) { return f(b, e, init); }
}
