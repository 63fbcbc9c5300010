// refstub (TEST INFRASTRUCTURE ONLY): g2o types named by Converter.h declarations.
#ifndef REFSTUB_G2O_SIX_DOF
#define REFSTUB_G2O_SIX_DOF
namespace g2o { class SE3Quat {}; }
#endif
