// ref_runtime.h — harness-side services of ref_runtime.cpp.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
#ifndef REF_RUNTIME_H
#define REF_RUNTIME_H
#include <functional>
#include "glsl_shim.h"
namespace glsl {
// runs entry() once per invocation id, the 32 invocations in lock-step wherever they meet at ballotARB
void run_subgroup(void (*entry)(), const uvec3 ids[32]);
void parallel_for(int64_t n, const std::function<void(int64_t)>& body);
void set_threads(int n);
}
#endif
