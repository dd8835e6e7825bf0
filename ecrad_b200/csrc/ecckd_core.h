// ecckd_core.h -- ecCKD gas-optics model and generalised cloud optics: device-side descriptors.
//
// Reference map:  CkdModel / CkdGas  <- radiation/radiation_ecckd.F90:34-118 (ckd_model_type), radiation_ecckd_gas.F90:38-72
//                 GcoType            <- radiation/radiation_general_cloud_optics_data.F90:33-62
// All look-up tables live in one device array of doubles ("ckdtab"), g-point fastest exactly as the reference stores them
// (molar_abs(ng, npress, ntemp [, nconc]), planck_function(ng, nplanck), mass_ext(ng, nre)), so lanes = g-points read
// contiguous memory; the descriptors hold element offsets into it.
#pragma once
#include <stddef.h>

#include "hd.h"

namespace ecb {

enum { CKD_MAXGAS = 12 };
enum { CKD_CONC_NONE = 0, CKD_CONC_LINEAR = 1, CKD_CONC_LUT = 2, CKD_CONC_RELATIVE_LINEAR = 3 };   // radiation_ecckd_gas.F90:27-32

struct CkdGas {
  int dep;            // concentration dependence
  int slot;           // index into DevIn::gas[] of this gas' mole-fraction array, -1: composite / not provided (zero)
  int n_mole_frac;
  int lut;            // index among the look-up-table gases of the model (dep == CKD_CONC_LUT), else -1
  double reference_mole_frac, log_mole_frac1, d_log_mole_frac;
  double mole_frac1;  // exp(log_mole_frac1), evaluated once on the host (radiation_ecckd.F90:586)
  double mmr_scaling; // local_concentration_scaling of this gas when the gas arrays are mass mixing ratios (mixed gas models)
  size_t off;         // molar_abs
};

struct CkdModel {
  int ng, npress, ntemp, nplanck, ngas, nlut;
  double log_pressure1, d_log_pressure, d_temperature, temperature1_planck, d_temperature_planck;
  size_t off_temperature1, off_planck, off_solar, off_rayleigh;
  long long off_solar_amp;   // norm_amplitude_solar_irradiance (read_spectral_solar_cycle, radiation_ecckd.F90:421-431), -1 if not registered
  CkdGas gas[CKD_MAXGAS];
};

struct GcoType { int nre; double re0, dre; size_t off_me, off_ssa, off_g; };

struct CkdMeta {
  CkdModel lw, sw;
  GcoType gco_lw[2], gco_sw[2];   // cloud types: 0 liquid, 1 ice
};

}  // namespace ecb
