// Stand-in header (oracle build only): everything lives in the master shim header.
#include <boost/numeric/odeint.hpp>
