// Build scaffolding for oracle/_ref ONLY (test infrastructure, never part of the product).
// Force-included ahead of the reference's sources so that its MSVC-isms compile with g++:
//  * std::exception("msg") (MSVC extension, ~150 call sites) -> a runtime_error-derived type;
//  * unqualified isinf/isnan (Functions.h:67,73; Dual.h:333,342);
//  * ctime_s (Utilities.cpp:344).
#pragma once
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>
#include <functional>
#include <memory>
#include <algorithm>
#include <numeric>
#include <random>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <iostream>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <queue>
#include <barrier>
#include <ranges>
#include <cstring>
#include <cmath>
#include <atomic>
#include <array>
#include <tuple>
#include <chrono>
#include <regex>
#include <map>
#include <iterator>
#include <limits>
#include <immintrin.h>
namespace std { struct dl_exception : runtime_error { dl_exception(const char* m = "") : runtime_error(m) {} }; }
#define exception dl_exception
using std::isinf; using std::isnan;
#include <ctime>
inline int ctime_s(char* buf, size_t n, const time_t* t) { (void)n; return ctime_r(t, buf) ? 0 : 1; }
