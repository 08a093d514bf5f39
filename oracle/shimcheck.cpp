// Compile-check of the product's C++ drop-in headers (include/libcluster.h, include/distributions.h)
// against the Eigen stand-in: the user code of the reference's test/cluster_test.cpp:38-66 must compile.
#include "distributions.h"
#include "libcluster.h"

using namespace Eigen;
using namespace libcluster;
using namespace distributions;

double use_grouped(const vMatrixXd& X) {
  std::vector<GDirichlet> weights;
  std::vector<GaussWish> clusters;
  vMatrixXd qZ;
  double F = learnGMC(X, qZ, weights, clusters, PRIORVAL, -1, false, true);
  for (std::vector<GDirichlet>::iterator j = weights.begin(); j < weights.end(); ++j) std::cout << j->Elogweight().exp().transpose();
  for (std::vector<GaussWish>::iterator k = clusters.begin(); k < clusters.end(); ++k) std::cout << k->getmean() << k->getcov();
  return F;
}
double use_flat(const MatrixXd& X) {
  MatrixXd qZ;
  StickBreak sb(2.0);
  Dirichlet dir;
  std::vector<GaussWish> gw;
  std::vector<NormGamma> ng;
  double F = learnVDP(X, qZ, sb, gw) + learnBGMM(X, qZ, dir, gw, 1.0, 5) + learnDGMM(X, qZ, dir, ng, 1.0, -1, false, 4);
  GaussWish c(1.0, (unsigned)X.cols());
  c.addobs(VectorXd::Ones(X.rows()), X);
  c.update();
  VectorXd e = c.Eloglike(X);
  ArrayXb s = c.splitobs(X);
  return F + e.sum() + c.fenergy() + c.getN() + c.getprior() + (double)s.count() + dir.fenergy() + sb.getNk().sum();
}
int main() { return 0; }
