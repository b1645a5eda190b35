// Source-compatibility forwarder: <RLGymSim_CPP/Utils/StateSetters/StateSetter.h> of the reference resolves to the B200 shim (include/rlgym_b200_shim.hpp).
#pragma once
#include "../../../../rlgym_b200_shim.hpp"
