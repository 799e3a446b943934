/* A plain C client of the drop-in boundary: includes the public header, loads the shared library the way a
 * foreign-function binding would (dlopen / dlsym, no C++ and no torch in sight) and checks the contract that can be
 * checked without a GPU: ABI version, diagnostics, argument validation (TSDE_EINVAL) and the no-op on empty launches.
 * Built and run by tests/test_host_logic.py::test_c_client_of_the_abi. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "torchsde_b200.h"

typedef int (*abi_version_fn)(void);
typedef const char* (*error_string_fn)(int);
typedef int64_t (*kernel_launches_fn)(int32_t);
typedef int (*step_euler_fn)(const tsde_launch*, const tsde_noise*, const void*, const void*, const void*, double, void*);
typedef int (*linear_interp_fn)(const tsde_launch*, const void*, const void*, double, double, void*);

#define REQUIRE(cond)                                                  \
  do {                                                                 \
    if (!(cond)) {                                                     \
      fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond);        \
      return 1;                                                        \
    }                                                                  \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 3;
  }
  abi_version_fn abi_version = (abi_version_fn)dlsym(lib, "tsde_abi_version");
  error_string_fn error_string = (error_string_fn)dlsym(lib, "tsde_error_string");
  kernel_launches_fn kernel_launches = (kernel_launches_fn)dlsym(lib, "tsde_kernel_launches");
  step_euler_fn step_euler = (step_euler_fn)dlsym(lib, "tsde_step_euler");
  linear_interp_fn linear_interp = (linear_interp_fn)dlsym(lib, "tsde_linear_interp");
  REQUIRE(abi_version && error_string && kernel_launches && step_euler && linear_interp);
  REQUIRE(abi_version() == TSDE_ABI_VERSION);
  REQUIRE(strstr(error_string(TSDE_EINVAL), "invalid argument") != NULL);
  REQUIRE(kernel_launches(TSDE_KERNEL_GEN_CTA) == 0 && kernel_launches(TSDE_KERNEL_GEN_TMA) == 0);
  REQUIRE(kernel_launches(42) == -1);

  float buffer[64] = {0};
  tsde_launch L;
  memset(&L, 0, sizeof L);
  L.dtype = TSDE_F32;
  L.noise_type = TSDE_NOISE_DIAGONAL;
  L.rows = 0; /* empty launch */
  L.d = 8;
  L.m = 8;
  tsde_noise nz;
  memset(&nz, 0, sizeof nz);
  nz.source = TSDE_SRC_MEMORY;
  nz.w = buffer;
  REQUIRE(sizeof(tsde_launch) == 40 && sizeof(tsde_noise) == 80);
  REQUIRE(step_euler(&L, &nz, buffer, buffer, buffer, 0.1, buffer) == 0);       /* nothing to do */
  REQUIRE(linear_interp(&L, buffer, buffer, 0.5, 0.5, buffer) == 0);
  REQUIRE(step_euler(NULL, &nz, buffer, buffer, buffer, 0.1, buffer) == TSDE_EINVAL);
  L.rows = 4;
  REQUIRE(step_euler(&L, NULL, buffer, buffer, buffer, 0.1, buffer) == TSDE_EINVAL);
  REQUIRE(step_euler(&L, &nz, NULL, buffer, buffer, 0.1, buffer) == TSDE_EINVAL);
  L.m = 4; /* diagonal noise needs m == d */
  REQUIRE(step_euler(&L, &nz, buffer, buffer, buffer, 0.1, buffer) == TSDE_EINVAL);
  L.m = 8;
  L.dtype = 7;
  REQUIRE(step_euler(&L, &nz, buffer, buffer, buffer, 0.1, buffer) == TSDE_EINVAL);
  printf("c client ok, abi %d\n", abi_version());
  dlclose(lib);
  return 0;
}
