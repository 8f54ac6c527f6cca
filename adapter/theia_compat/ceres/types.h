// Shim for <ceres/types.h>: the three enums src/theia/sfm/bundle_adjustment/bundle_adjustment.h:38 pulls in,
// with upstream Ceres' numeric values, for builds where Ceres is absent (it is absent in this image).
#ifndef THEIA_COMPAT_CERES_TYPES_H_
#define THEIA_COMPAT_CERES_TYPES_H_
namespace ceres {
enum LinearSolverType { DENSE_NORMAL_CHOLESKY = 0, DENSE_QR = 1, SPARSE_NORMAL_CHOLESKY = 2, DENSE_SCHUR = 3, SPARSE_SCHUR = 4,
                        ITERATIVE_SCHUR = 5, CGNR = 6 };
enum PreconditionerType { IDENTITY = 0, JACOBI = 1, SCHUR_JACOBI = 2, CLUSTER_JACOBI = 3, CLUSTER_TRIDIAGONAL = 4 };
enum VisibilityClusteringType { CANONICAL_VIEWS = 0, SINGLE_LINKAGE = 1 };
}  // namespace ceres
#endif
