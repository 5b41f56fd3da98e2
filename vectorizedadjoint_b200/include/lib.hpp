// lib.hpp -- umbrella header of the B200-native drop-in for the reference's lib/include/lib.hpp.
#include "runge_kutta.hpp"
#include "backpropagation.hpp"
