// forwarding header: the reference declares UniformParameter here (src/include/UniformParameter.h)
#include "DistributionParameters.h"
