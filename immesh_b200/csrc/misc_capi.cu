// immesh_b200 -- version / error string entry points.
#include "common_host.hpp"
extern "C" {
const char* immesh_last_error(void) { return immesh::last_error_storage().c_str(); }
const char* immesh_version(void) { return "immesh_b200 0.1.0 (sm_100a)"; }
}
