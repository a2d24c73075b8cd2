// The reference configures the model with compile-time macros (parameters.hpp:1-15).  Here the
// model is a runtime struct (b200_edm_model); only the two names that user code such as
// Driver.cu refers to are kept, with the reference's values.
#ifndef B200_COMPAT_PARAMETERS_HPP
#define B200_COMPAT_PARAMETERS_HPP
#define noSpikes 3
#define timeHorizon 5.0f
#endif
