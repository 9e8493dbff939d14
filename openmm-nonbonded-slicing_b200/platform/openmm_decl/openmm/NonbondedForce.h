#ifndef NBS_DECL_NONBONDEDFORCE_H_
#define NBS_DECL_NONBONDEDFORCE_H_
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class Context;
class ForceImpl;
class NonbondedForce {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2, Ewald = 3, PME = 4, LJPME = 5 };
    NonbondedForce();
    virtual ~NonbondedForce();
    int getNumParticles() const;
    int getNumExceptions() const;
    int getNumGlobalParameters() const;
    int getNumParticleParameterOffsets() const;
    int getNumExceptionParameterOffsets() const;
    NonbondedMethod getNonbondedMethod() const;
    double getCutoffDistance() const;
    bool getUseSwitchingFunction() const;
    double getSwitchingDistance() const;
    double getReactionFieldDielectric() const;
    bool getUseDispersionCorrection() const;
    bool getExceptionsUsePeriodicBoundaryConditions() const;
    void getParticleParameters(int index, double& charge, double& sigma, double& epsilon) const;
    void getExceptionParameters(int index, int& particle1, int& particle2, double& chargeProd, double& sigma, double& epsilon) const;
    const std::string& getGlobalParameterName(int index) const;
    double getGlobalParameterDefaultValue(int index) const;
    void getParticleParameterOffset(int index, std::string& parameter, int& particleIndex, double& chargeScale, double& sigmaScale, double& epsilonScale) const;
    void getExceptionParameterOffset(int index, std::string& parameter, int& exceptionIndex, double& chargeProdScale, double& sigmaScale, double& epsilonScale) const;
protected:
    virtual ForceImpl* createImpl() const;
};
}
#endif
