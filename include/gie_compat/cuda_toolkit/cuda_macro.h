// Stand-in for the reference's include/cuda_toolkit/cuda_macro.h: vector types and the error convention.
// The reference exit(1)s on a CUDA error (cuda_macro.h:21-31); the wrappers in this tree throw gie::Error instead
// (define GIE_COMPAT_EXIT_ON_ERROR to get the reference's behaviour).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector_types.h>
#include <vector_functions.h>
#include "gie_b200.h"

namespace gie {
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};
inline void check(int status, const char *where)
{
    if (status == GIE_OK) return;
    std::string msg = std::string(where) + ": gie status " + std::to_string(status) + ": " + gie_last_error();
#ifdef GIE_COMPAT_EXIT_ON_ERROR
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(1);
#else
    throw Error(status, msg);
#endif
}
}  // namespace gie
#define GIE_CHECK(call) ::gie::check((call), #call)
