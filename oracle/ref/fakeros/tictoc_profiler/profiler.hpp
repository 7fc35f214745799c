/* oracle/ref/fakeros -- CPU ORACLE, TEST INFRASTRUCTURE ONLY: the profiler the reference brackets its stages with; a no-op here. */
#pragma once
#include <string>
namespace ca {
struct Profiler {
    static void enable() {}
    static void tictoc(const std::string &) {}
    static void print_aggregated(std::ostream &) {}
};
}  // namespace ca
