// oracle/_ref/taesoo_ref -- command-line harness around the REFERENCE's own taesooLib BaseLib, compiled from the sources where
// they lie under /root/reference/work/taesooLib/BaseLib (see Makefile).  Test infrastructure only: tests/tools/make_golden_ref.py
// runs it to record what the reference computes (tests/golden/ref_*.npz); nothing in the product links or executes it.
//
// This file is original glue code (argument parsing, text I/O); every number it prints is produced by reference code:
//   fk       BoneForwardKinematics::setPoseDOF / global   (BaseLib/motion/BoneKinematics.cpp)
//            IK_sdls::LoaderToTree (quaternion root): setPoseDOF, setDQ, globalFrame, calcCOM, calcMomentumCOM
//                                                         (BaseLib/motion/IK_sdls/NodeWrap.cpp, Node.cpp, Liegroup.cpp)
//   euler    IK_sdls::LoaderToTree (Euler root, as ShortTrajOpt builds it): setQ, setDQ, globalFrame, calcCOM, calcMomentumCOM,
//            calcCOMjacobianTranspose, calcMomentumJacobianTranspose, calcJacobianTransposeAt, calcRotJacobianTranspose,
//            updateGrad_S_JT
//   quat     quater / vector3 / transf / Liegroup primitives  (BaseLib/math/quater.cpp, vector3.cpp, transf.cpp, Liegroup.cpp)
//   terrain  OBJloader::Terrain::height                       (BaseLib/motion/Terrain.cpp:305-393)
//   dof      MotionDOFcontainer / BinaryFile readers          (BaseLib/motion/MotionDOF.cpp, utility/tfile.cpp)
//   signal   Filter::GetGaussFilter / gaussFilter (BaseLib/math/Filter.cpp:25-31,257-277,347-389,569-630), NonuniformSpline with
//            zero end accelerations: getCurve / getSecondDeriv (BaseLib/math/BSpline.cpp:143-414), matrixn::sampleRow
//            (BaseLib/math/matrixn.cpp:781-805) -- the signal operators under fc.CDMTraj (gym_cdm2/module/CDMTraj.lua:709-775,875)
//   samplepose  MotionDOF::samplePose / MotionDOFinfo::blend (BaseLib/motion/MotionDOF.cpp:237-257,593-620)
//   hull     GrahamScan add_point / partition_points / build_hull / get_hull   (PhysicsLib/convexhull/graham.cpp), the hull
//            under math.projectTrajectoryToItsHull (gym_cdm2/ConvexHull2D.lua:150-186,362)
// Numbers travel as whitespace-separated %.17g text (round-trip exact for doubles).
#include "stdafx.h"
#include "motion/Mesh.h"
#define class struct               // Terrain::_tileAlongZ is a default-private member; the loader's seam blend (m::c1stitch) needs
#include "motion/Terrain.h"        // Eigen and cannot run here, the query can: the harness sets the flag after construction
#undef class
#include "motion/VRMLloader.h"
#include "motion/BoneKinematics.h"
#include "motion/MotionDOF.h"
#include "motion/MotionWrap.h"
#include "motion/Liegroup.h"
#include "motion/IK_sdls/NodeWrap.h"
#include "math/Operator.h"
#include "math/Filter.h"
#include "math/BSpline.h"
#include "../PhysicsLib/convexhull/graham.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

void __noop(...) {}                 // BaseLib/stdafx.cpp:5 (the TRACE sink of release builds; stdafx.cpp itself is not among the compiled files)

static std::vector<double> read_all(const char* path) {
  std::vector<double> v;
  FILE* f = fopen(path, "r");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  double d;
  while (fscanf(f, "%lf", &d) == 1) v.push_back(d);
  fclose(f);
  return v;
}
struct Out {
  FILE* f;
  explicit Out(const char* path) { f = fopen(path, "w"); if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); } }
  ~Out() { fclose(f); }
  void d(double x) { fprintf(f, "%.17g\n", x); }
  void v3(vector3 const& v) { d(v.x); d(v.y); d(v.z); }
  void q(quater const& r) { d(r.w); d(r.x); d(r.y); d(r.z); }
  void tf(transf const& t) { v3(t.translation); q(t.rotation); }
  void mat(matrixn const& m) { d(m.rows()); d(m.cols()); for (int i = 0; i < m.rows(); i++) for (int j = 0; j < m.cols(); j++) d(m(i, j)); }
  void vec(vectorn const& v) { d(v.size()); for (int i = 0; i < v.size(); i++) d(v[i]); }
};

static int cmd_fk(int argc, char** argv) {
  VRMLloader l(argv[2]);
  std::vector<double> in = read_all(argv[3]);
  Out out(argv[4]);
  const int nd = l.dofInfo.numDOF();                 // 60
  const int ndq = nd - 1;
  const int nb = l.numBone();
  const int row = nd + ndq;
  const int n = (int)in.size() / row;
  BoneForwardKinematics fk(&l);
  fk.init();
  IK_sdls::LoaderToTree tree(l, false, false);
  out.d(n); out.d(nb - 1); out.d(nd);
  for (int r = 0; r < n; r++) {
    vectorn pose(nd), dq(ndq);
    for (int i = 0; i < nd; i++) pose[i] = in[(size_t)r * row + i];
    for (int i = 0; i < ndq; i++) dq[i] = in[(size_t)r * row + nd + i];
    fk.setPoseDOF(pose);
    for (int b = 1; b < nb; b++) out.tf(fk.global(b));
    tree.setPoseDOF(l.dofInfo, pose);
    tree.setDQ(dq);
    for (int b = 1; b < nb; b++) out.tf(tree.globalFrame(b));
    out.v3(tree.calcCOM(l));
    Liegroup::dse3 h = tree.calcMomentumCOM(l);
    for (int i = 0; i < 6; i++) out.d(h[i]);
    for (int b = 1; b < nb; b++) { out.v3(tree.getWorldAngVel(b)); out.v3(tree.getWorldVelocity(b)); }
  }
  return 0;
}

// argv: euler <wrl> <in> <out> <nmarker> then per marker: bone lx ly lz
static int cmd_euler(int argc, char** argv) {
  VRMLloader l(argv[2]);
  std::vector<double> in = read_all(argv[3]);
  Out out(argv[4]);
  const int nm = atoi(argv[5]);
  std::vector<int> mb(nm); std::vector<vector3> ml(nm);
  for (int k = 0; k < nm; k++) { mb[k] = atoi(argv[6 + 4 * k]); ml[k] = vector3(atof(argv[7 + 4 * k]), atof(argv[8 + 4 * k]), atof(argv[9 + 4 * k])); }
  IK_sdls::LoaderToTree tree(l, true, false);
  const int nq = tree.nDOF();
  const int nb = l.numBone();
  const int row = 2 * nq + 3;                        // q, dq, deltaS (for updateGrad_S_JT)
  const int n = (int)in.size() / row;
  out.d(n); out.d(nb - 1); out.d(nq); out.d(nm);
  for (int r = 0; r < n; r++) {
    vectorn q(nq), dq(nq);
    for (int i = 0; i < nq; i++) { q[i] = in[(size_t)r * row + i]; dq[i] = in[(size_t)r * row + nq + i]; }
    vector3 dS(in[(size_t)r * row + 2 * nq], in[(size_t)r * row + 2 * nq + 1], in[(size_t)r * row + 2 * nq + 2]);
    tree.setQ(q);
    tree.setDQ(dq);
    for (int b = 1; b < nb; b++) out.tf(tree.globalFrame(b));
    out.v3(tree.calcCOM(l));
    Liegroup::dse3 h = tree.calcMomentumCOM(l);
    for (int i = 0; i < 6; i++) out.d(h[i]);
    matrixn JT;
    tree.calcCOMjacobianTranspose(l, JT); out.mat(JT);
    tree.calcMomentumJacobianTranspose(l, JT); out.mat(JT);
    for (int k = 0; k < nm; k++) {
      out.v3(tree.globalFrame(mb[k]) * ml[k]);
      tree.calcJacobianTransposeAt(JT, mb[k], ml[k]); out.mat(JT);
      tree.calcRotJacobianTranspose(JT, mb[k]); out.mat(JT);
      vectorn g(nq); g.setAllValue(0.0);
      tree.updateGrad_S_JT(g, dS, mb[k], ml[k]); out.vec(g);
    }
    vectorn back(nq + 1); back.setAllValue(0.0);
    tree.getQ(back); for (int i = 0; i < nq; i++) out.d(back[i]);
    tree.getDQ(back); for (int i = 0; i < nq; i++) out.d(back[i]);
  }
  return 0;
}

// quat <in> <out>: rows of 12 doubles  [q1 wxyz | q2 wxyz | v xyz | t]
static int cmd_quat(int argc, char** argv) {
  std::vector<double> in = read_all(argv[2]);
  Out out(argv[3]);
  const int row = 12, n = (int)in.size() / row;
  out.d(n);
  const char* chans[4] = {"YZX", "XYZ", "ZXY", "X"};
  for (int r = 0; r < n; r++) {
    const double* a = &in[(size_t)r * row];
    quater q1(a[0], a[1], a[2], a[3]), q2(a[4], a[5], a[6], a[7]);
    vector3 v(a[8], a[9], a[10]);
    double t = a[11];
    q1.normalize(); q2.normalize();
    out.q(q1); out.q(q2);
    quater m; m.mult(q1, q2); out.q(m);                                  // quater::mult
    out.q(q1.inverse());
    vector3 rv; rv.rotate(q1, v); out.v3(rv);                            // vector3::rotate
    for (int c = 0; c < 4; c++) { quater e; m_real ang[3] = {v.x, v.y, v.z}; e.setRotation(chans[c], ang); out.q(e); }   // setRotation(channels)
    quater ax; ax.setRotation(v, t); out.q(ax);                          // setRotation(axis, angle) (axis as given)
    quater ex; ex.exp(v); out.q(ex);                                     // quater::exp
    out.v3(q1.log());                                                    // quater::log
    out.v3(q1.rotationVector());
    out.d(q1.rotationAngle());
    quater dq; dq.difference(q1, q2); out.q(dq);
    quater sl; sl.safeSlerp(q1, q2, t); out.q(sl);
    quater rotY, offset; q1.decompose(rotY, offset); out.q(rotY); out.q(offset);                   // rotationY * offset
    quater nt, tw; q1.decomposeNoTwistTimesTwist(vector3(0, 1, 0), nt, tw); out.q(nt); out.q(tw);
    quater aa; aa.axisToAxis(vector3(0, 1, 0), rv); out.q(aa);
    matrix4 M; M.setRotation(q1); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out.d(M.m[i][j]);
    quater qm; qm.setRotation(M); out.q(qm);                             // quater::setRotation(matrix4)
    // transf and Liegroup
    transf T1(q1, v), T2(q2, vector3(a[9], a[10], a[8]));
    transf T12; T12.mult(T1, T2); out.tf(T12);
    out.tf(T1.inverse());
    out.v3(T1 * v);
    Liegroup::se3 tw12 = Liegroup::twist(T1, T2, 1.0 / 30.0); for (int i = 0; i < 6; i++) out.d(tw12[i]);   // transf::twist
    Liegroup::se3 V(a[0], a[5], a[10], a[3], a[8], a[1]);
    matrix3 I3; I3.setValue(1.0 + a[0] * a[0], 0.1, 0.2, 0.1, 2.0 + a[1] * a[1], 0.3, 0.2, 0.3, 1.5 + a[2] * a[2]);
    Liegroup::Inertia I(a[11] + 2.0, I3, (a[11] + 2.0) * v);
    Liegroup::dse3 IV = I * V; for (int i = 0; i < 6; i++) out.d(IV[i]);
    Liegroup::dse3 d1 = IV.dAd(T1); for (int i = 0; i < 6; i++) out.d(d1[i]);
    Liegroup::dse3 d2 = IV.inv_dAd(T1); for (int i = 0; i < 6; i++) out.d(d2[i]);
    Liegroup::se3 lg; lg.log(T1); for (int i = 0; i < 6; i++) out.d(lg[i]);
    transf Te = V.exp(); out.tf(Te);
    out.d(sop::map(t, 0.2, 0.9, -3.0, 5.0)); out.d(sop::clampMap(t * 2, 0.2, 0.9, -3.0, 5.0));
  }
  return 0;
}

// terrain <in> <out>: in = rows cols sizeX sizeZ heightMax tile | image (rows*cols uint16 values) | nq | nq * (x z)
static int cmd_terrain(int argc, char** argv) {
  std::vector<double> in = read_all(argv[2]);
  Out out(argv[3]);
  size_t p = 0;
  const int rows = (int)in[p++], cols = (int)in[p++];
  const double sx = in[p++], sz = in[p++], hmax = in[p++];
  const int tile = (int)in[p++];
  OBJloader::Raw2Bytes image;
  image.setSize(rows, cols);
  for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) image(y, x) = (unsigned short)in[p++];
  OBJloader::Terrain T(&image, sx, sz, hmax, 1, 1, false);
  T._tileAlongZ = tile != 0;
  const int nq = (int)in[p++];
  out.d(nq);
  out.mat(T.getHeightMap());
  for (int k = 0; k < nq; k++) {
    vector2 x(in[p], in[p + 1]); p += 2;
    vector3 normal;
    double h = T.height(x);                    // the overload gym_cdm2 calls (deepmimiccdm-v1.lua:415-424)
    double h2 = T.height(x, normal);           // overload with the surface normal (no bounds test)
    out.d(h); out.d(h2); out.v3(normal);
  }
  return 0;
}

// dof <file.dof> <out>: MotionDOFcontainer -> frames x dofs matrix
static int cmd_dof(int argc, char** argv) {
  VRMLloader l(argv[2]);
  MotionDOFcontainer c(l.dofInfo, argv[3]);
  Out out(argv[4]);
  out.d(c.mot.numFrames()); out.d(c.mot.numDOF());
  for (int i = 0; i < c.mot.numFrames(); i++) for (int j = 0; j < c.mot.numDOF(); j++) out.d(c.mot(i, j));
  return 0;
}

// hull <in> <out>: in = sets of points: n, then n (x, y) pairs, repeated -> per set: closed hull polygon as a matrix
static int cmd_hull(int argc, char** argv) {
  std::vector<double> in = read_all(argv[2]);
  Out out(argv[3]);
  size_t p = 0;
  while (p < in.size()) {
    const int n = (int)in[p++];
    GrahamScan g;
    for (int i = 0; i < n; i++) { g.add_point(std::make_pair(in[p], in[p + 1])); p += 2; }
    g.partition_points();
    g.build_hull();
    matrixn h;
    g.get_hull(h);
    out.mat(h);
  }
  return 0;
}

// signal <in> <out>: in = rows cols | matrix | nk | kernel sizes | delta | nt | sample times
static int cmd_signal(int argc, char** argv) {
  std::vector<double> in = read_all(argv[2]);
  Out out(argv[3]);
  size_t p = 0;
  const int rows = (int)in[p++], cols = (int)in[p++];
  matrixn M(rows, cols);
  for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) M(i, j) = in[p++];
  const int nk = (int)in[p++];
  for (int k = 0; k < nk; k++) {
    const int ks = (int)in[p++];
    vectorn kernel;
    Filter::GetGaussFilter(ks, kernel);
    out.vec(kernel);
    matrixn F(M);
    Filter::gaussFilter(ks, F);                      // math.gaussFilter; math.filter is GetGaussFilter + LTIFilter(1, ...), the same two calls
    out.mat(F);
  }
  const int delta = (int)in[p++];
  intvectorn frames; frames.colon(0, rows, delta);   // CT.colon(0, rows, delta)
  vectorn timing(frames.size());
  matrixn control(frames.size(), cols);
  for (int i = 0; i < frames.size(); i++) { timing[i] = frames[i]; for (int j = 0; j < cols; j++) control(i, j) = M(frames[i], j); }
  NonuniformSpline spl(timing, control);
  const int nf = (int)timing[timing.size() - 1] + 1;
  vectorn nt; nt.linspace(timing[0], nf - 1, nf);
  matrixn curve, acc;
  spl.getCurve(nt, curve); spl.getSecondDeriv(nt, acc);
  out.vec(timing); out.mat(curve); out.mat(acc);
  const int ns = (int)in[p++];
  out.d(ns);
  for (int k = 0; k < ns; k++) { vectorn r; M.sampleRow(in[p++], r); out.vec(r); }
  return 0;
}

// samplepose <wrl> <file.dof> <in> <out>: in = sample times -> MotionDOF::samplePose rows
static int cmd_samplepose(int argc, char** argv) {
  VRMLloader l(argv[2]);
  MotionDOFcontainer c(l.dofInfo, argv[3]);
  std::vector<double> in = read_all(argv[4]);
  Out out(argv[5]);
  out.d((double)in.size()); out.d(c.mot.numDOF());
  for (size_t k = 0; k < in.size(); k++) {
    vectorn pose;
    c.mot.samplePose(in[k], pose);
    for (int j = 0; j < pose.size(); j++) out.d(pose[j]);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: taesoo_ref fk|euler|quat|terrain|dof|hull|signal|samplepose ...\n"); return 1; }
  std::string c = argv[1];
  if (c == "fk" && argc >= 5) return cmd_fk(argc, argv);
  if (c == "euler" && argc >= 6) return cmd_euler(argc, argv);
  if (c == "quat" && argc >= 4) return cmd_quat(argc, argv);
  if (c == "terrain" && argc >= 4) return cmd_terrain(argc, argv);
  if (c == "dof" && argc >= 5) return cmd_dof(argc, argv);
  if (c == "hull" && argc >= 4) return cmd_hull(argc, argv);
  if (c == "signal" && argc >= 4) return cmd_signal(argc, argv);
  if (c == "samplepose" && argc >= 6) return cmd_samplepose(argc, argv);
  fprintf(stderr, "bad command line\n");
  return 1;
}
