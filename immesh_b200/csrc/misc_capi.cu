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
    if (on == 2) immesh::profiler().start_timeline();   // also record every launch's start / end time
    else immesh::profiler().timeline_on = false;
    return IMMESH_OK;
}
int immesh_profile_reset(void) {
    immesh::profiler().totals.clear();
    immesh::profiler().launches = 0;
    return IMMESH_OK;
}
long long immesh_launch_count(void) { return immesh::profiler().launches; }
// "name t0_ms t1_ms\n" per launch recorded since immesh_profile_enable(2); returns the number of bytes needed
int immesh_profile_timeline(char* buf, int cap) {
    std::string s;
    for (auto& sp : immesh::profiler().timeline) {
        char line[256];
        std::snprintf(line, sizeof(line), "%s %.6f %.6f\n", sp.name, sp.t0, sp.t1);
        s += line;
    }
    if (buf && cap > 0) {
        std::strncpy(buf, s.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)s.size() + 1;
}
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
