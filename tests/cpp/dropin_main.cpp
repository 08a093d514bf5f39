// A libcluster user program written against the reference's API (cf. test/cluster_test.cpp:38-66),
// built with this repo's include/libcluster.h + include/distributions.h and linked to liblcb200.so.
// Reads the observations from a binary file so that it can run where /root/reference is absent.
//   dropin <file> : file = int32 J, int32 D, then per group int32 N_j and N_j*D doubles (row-major)
#include <cstdio>
#include <cstdlib>
#include <iomanip>

#include "distributions.h"
#include "libcluster.h"

using namespace std;
using namespace Eigen;
using namespace libcluster;
using namespace distributions;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int J = 0, D = 0;
  if (fread(&J, 4, 1, f) != 1 || fread(&D, 4, 1, f) != 1) return 2;
  vMatrixXd X(J);
  int total = 0;
  for (int j = 0; j < J; ++j) {
    int N = 0;
    if (fread(&N, 4, 1, f) != 1) return 2;
    X[j] = MatrixXd(N, D);
    vector<double> buf((size_t)N * D);
    if (N && fread(buf.data(), 8, buf.size(), f) != buf.size()) return 2;
    for (int n = 0; n < N; ++n)
      for (int d = 0; d < D; ++d) X[j](n, d) = buf[(size_t)n * D + d];
    total += N;
  }
  fclose(f);
  cout << setprecision(15);
  try {
    // grouped model, exactly the call of test/cluster_test.cpp:51
    vector<GDirichlet> weights;
    vector<GaussWish> clusters;
    vMatrixXd qZgroup;
    double F = learnGMC(X, qZgroup, weights, clusters, PRIORVAL, -1, false, false);
    cout << "GMC F " << F << " K " << clusters.size() << endl;
    for (size_t k = 0; k < clusters.size(); ++k) cout << "mean " << clusters[k].getmean() << endl;
    cout << "w0 " << weights[0].Elogweight().exp().transpose() << endl;
    // flat model on the concatenation
    MatrixXd Xcat(total, D);
    for (int j = 0, r = 0; j < J; ++j)
      for (Index n = 0; n < X[j].rows(); ++n, ++r)
        for (int d = 0; d < D; ++d) Xcat(r, d) = X[j](n, d);
    MatrixXd qZ;
    Dirichlet dir;
    vector<GaussWish> cl2;
    double F2 = learnBGMM(Xcat, qZ, dir, cl2, PRIORVAL, -1, false, 1);
    double rowsum = 0;
    for (Index n = 0; n < qZ.rows(); ++n) {
      double s = 0;
      for (Index k = 0; k < qZ.cols(); ++k) s += qZ(n, k);
      rowsum += s;
    }
    cout << "BGMM F " << F2 << " K " << cl2.size() << " rowsum " << rowsum << endl;
    // caller-supplied weight priors are kept by the fit (src/cluster.cpp:653,684)
    {
      Dirichlet dir5(5.0);
      vector<GaussWish> cl3;
      MatrixXd q3;
      double F3 = learnBGMM(Xcat, q3, dir5, cl3, PRIORVAL, -1, false, 1);
      cout << "BGMM5 F " << F3 << " K " << cl3.size() << " Elogw " << dir5.Elogweight().transpose() << endl;
      StickBreak sb3(3.0);
      vector<GaussWish> cl4;
      MatrixXd q4;
      double F4 = learnVDP(Xcat, q4, sb3, cl4, PRIORVAL, -1, false, 1);
      cout << "VDP3 F " << F4 << " K " << cl4.size() << " Elogw " << sb3.Elogweight().transpose() << endl;
    }
    // operator surface
    GaussWish c(PRIORVAL, (unsigned)D);
    c.addobs(VectorXd::Ones(Xcat.rows()), Xcat);
    c.update();
    cout << "OPS N " << c.getN() << " Esum " << c.Eloglike(Xcat).sum() << " F " << c.fenergy() << endl;
    // error behaviour: nthreads < 1 -> std::invalid_argument (cluster.cpp:576-577)
    try {
      learnBGMM(Xcat, qZ, dir, cl2, PRIORVAL, -1, false, 0);
      cout << "ERR none" << endl;
    } catch (const invalid_argument& e) {
      cout << "ERR invalid_argument " << e.what() << endl;
    }
  } catch (const exception& e) {
    cout << "EXCEPTION " << e.what() << endl;
    return 1;
  }
  return 0;
}
