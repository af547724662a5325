// mphx_internal.h -- shared by the host-only and the CUDA translation units of libmphx.so
#ifndef MPHX_INTERNAL_H
#define MPHX_INTERNAL_H
#include <string>

namespace mphx {
void set_last_error(const std::string &msg);
}

#endif
