// immesh_b200 -- version / error string / profiling entry points.
#include <cstring>

#include "common_host.hpp"

namespace immesh {
Profiler& profiler() {
    static Profiler p;
    return p;
}
}  // namespace immesh

extern "C" {
const char* immesh_last_error(void) { return immesh::last_error_storage().c_str(); }
const char* immesh_version(void) { return "immesh_b200 0.1.0 (sm_100a)"; }

int immesh_profile_enable(int on) {
    immesh::profiler().enabled = on != 0;
    return IMMESH_OK;
}
int immesh_profile_reset(void) {
    immesh::profiler().totals.clear();
    immesh::profiler().launches = 0;
    return IMMESH_OK;
}
long long immesh_launch_count(void) { return immesh::profiler().launches; }
// writes "name ms launches\n" lines into buf; returns the number of bytes needed
int immesh_profile_report(char* buf, int cap) {
    std::string s;
    for (auto& kv : immesh::profiler().totals) {
        char line[256];
        std::snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
        s += line;
    }
    if (buf && cap > 0) {
        std::strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)s.size() + 1;
}
}
