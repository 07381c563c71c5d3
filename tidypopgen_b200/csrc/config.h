// Process-wide configuration (include/tpg_b200.h: tpg_config).  The job path reads cfg(); the
// TPG_* environment variables are read in ONE place, tpg_config_init (config.cpp).
#pragma once
#include "../../include/tpg_b200.h"

namespace tpg {
// Snapshot of the configuration in force (initialised from tpg_config_init on first use).
tpg_config cfg();
}  // namespace tpg
