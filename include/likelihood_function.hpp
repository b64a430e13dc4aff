// Same abstract interface as the reference's include/likelihood_function.hpp (:5-49): what its samplers
// (MetropolisHastings source/mcmc.cpp:523,594; MultiNest / PolyChord wrappers) call once per proposed point.
#ifndef COSMO_PP_LIKELIHOOD_FUNCTION_HPP
#define COSMO_PP_LIKELIHOOD_FUNCTION_HPP

namespace Math
{

class LikelihoodFunction
{
public:
    virtual ~LikelihoodFunction() {}
    // -2 ln(likelihood) at params[0 .. nParams-1]
    virtual double calculate(double* params, int nParams) = 0;
    // exact version where calculate() approximates; defaults to calculate()
    virtual double calculateExact(double* params, int nParams) { return calculate(params, nParams); }
};

class LikelihoodWithDerivs : public LikelihoodFunction
{
public:
    virtual ~LikelihoodWithDerivs() {}
    virtual double calculate(double* params, int nParams) = 0;
    // -2 d ln(likelihood) / d params[i]
    virtual double calculateDeriv(double* params, int nParams, int i) = 0;
};

} // namespace Math

#endif
