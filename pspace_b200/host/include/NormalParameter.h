// forwarding header: the reference declares NormalParameter here (src/include/NormalParameter.h)
#include "DistributionParameters.h"
