// forwarding header: the reference declares ExponentialParameter here (src/include/ExponentialParameter.h)
#include "DistributionParameters.h"
