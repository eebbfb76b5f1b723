// oracle/shim/ros/ros.h -- TEST INFRASTRUCTURE.  ros::Time as a deterministic step counter: every now() advances the clock
// by ros::shim_tick() seconds, so the reference's wall-clock budgets become exact iteration budgets (tick 1, limit N = N
// iterations; tick 0 = no time-out).  ROS_WARN / ROS_ERROR are counted, not printed.
#pragma once
#include <string>
namespace ros {
inline double &shim_clock() { static double t = 0.0; return t; }
inline double &shim_tick() { static double d = 0.0; return d; }
inline long long &shim_warnings() { static long long n = 0; return n; }
struct Duration { double s; Duration() : s(0) {} explicit Duration(double s_) : s(s_) {} double toSec() const { return s; } };
struct Time {
    double t;
    Time() : t(0) {}
    explicit Time(double t_) : t(t_) {}
    static Time now() { shim_clock() += shim_tick(); return Time(shim_clock()); }
    Duration operator-(const Time &o) const { return Duration(t - o.t); }
    Time operator+(const Duration &d) const { return Time(t + d.s); }
};
}
#define ROS_WARN(...) do { ros::shim_warnings()++; } while (0)
#define ROS_ERROR(...) do { ros::shim_warnings()++; } while (0)
#define ROS_INFO(...) do { } while (0)
