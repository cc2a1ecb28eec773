// extern "C" shim over the reference's FROST-generated Go1 kinematics, compiled from the sources
// where they lie under /root/reference (never copied).  Exposes the *_raw entry points declared in
// src/go1_example/include/Expressions/{FR,FL,RR,RL}_foot.hh:20 and J_{FR,FL,RR,RL}.hh:20.
// Output conventions (as generated): foot = 1x3; J = 3x22 column-major.
// Test infrastructure only.
namespace SymFunction {
void FR_foot_raw(double *, const double *);
void FL_foot_raw(double *, const double *);
void RR_foot_raw(double *, const double *);
void RL_foot_raw(double *, const double *);
void J_FR_raw(double *, const double *);
void J_FL_raw(double *, const double *);
void J_RR_raw(double *, const double *);
void J_RL_raw(double *, const double *);
}  // namespace SymFunction

extern "C" {
// leg: 0 FR, 1 FL, 2 RR, 3 RL (go1Sub.cpp:82 "Foot order")
void ref_go1_foot(int leg, const double *var22, double *p3) {
  switch (leg) {
    case 0: SymFunction::FR_foot_raw(p3, var22); break;
    case 1: SymFunction::FL_foot_raw(p3, var22); break;
    case 2: SymFunction::RR_foot_raw(p3, var22); break;
    default: SymFunction::RL_foot_raw(p3, var22); break;
  }
}
void ref_go1_jac(int leg, const double *var22, double *J3x22_colmajor) {
  switch (leg) {
    case 0: SymFunction::J_FR_raw(J3x22_colmajor, var22); break;
    case 1: SymFunction::J_FL_raw(J3x22_colmajor, var22); break;
    case 2: SymFunction::J_RR_raw(J3x22_colmajor, var22); break;
    default: SymFunction::J_RL_raw(J3x22_colmajor, var22); break;
  }
}
}
