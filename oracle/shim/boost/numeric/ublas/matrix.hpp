// Stand-in header (oracle build only).
#include <boost/numeric/ublas/shim_matrix.hpp>
