// refstub (TEST INFRASTRUCTURE ONLY): g2o types named by Converter.h declarations.
#ifndef REFSTUB_G2O_SEVEN_DOF
#define REFSTUB_G2O_SEVEN_DOF
namespace g2o { class Sim3 {}; }
#endif
