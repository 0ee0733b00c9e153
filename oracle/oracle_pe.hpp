// oracle/oracle_pe.hpp -- TEST INFRASTRUCTURE (see oracle_core.hpp). Paired-end restatement: placeholder.
#pragma once
#include "oracle_core.hpp"
namespace oracle {
inline bool run_pe(const Index&, const Params&, const char*, const char*, const char*, std::string&, Stats&) {
  fprintf(stderr, "oracle: paired-end restatement not built yet\n"); return false;
}
}
