// TEST INFRASTRUCTURE ONLY (oracle): forwards to the GSL stand-in.
#include "gsl_shim.h"
