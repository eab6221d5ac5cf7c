/* ORACLE / TEST INFRASTRUCTURE.  Stand-in for general/gr_prefs.h: no preference files exist here, every query
 * returns its default (digital_clock_recovery_mm_cc.cc:60 asks for "verbose"). */
#ifndef ORACLE_SHIM_GR_PREFS_H
#define ORACLE_SHIM_GR_PREFS_H
#include <string>
class gr_prefs {
 public:
  static gr_prefs* singleton() { static gr_prefs p; return &p; }
  bool get_bool(const std::string&, const std::string&, bool default_val) { return default_val; }
  long get_long(const std::string&, const std::string&, long default_val) { return default_val; }
  double get_double(const std::string&, const std::string&, double default_val) { return default_val; }
  const std::string get_string(const std::string&, const std::string&, const std::string& default_val) { return default_val; }
};
#endif
