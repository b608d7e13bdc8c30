// pb_host.h - host-side internals shared between the pipeline and helpers.
#pragma once
#include <stddef.h>
#include <stdint.h>

void pb_lapack_set_path(const char *path);
const char *pb_lapack_source();
void pb_lapack_allow_jacobi(bool on);
void pb_lapack_use_host(bool on); // true: call a run-time resolved dsyev_ instead of the built-in restatement
bool pb_eigen_solve3(double a[9], double w[3]);
