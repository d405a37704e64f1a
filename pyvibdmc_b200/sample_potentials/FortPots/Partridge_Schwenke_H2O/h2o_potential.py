"""water_pot(cds): the Partridge-Schwenke H2O surface on the GPU.  Replaces the reference's f2py
wrapper (FortPots/Partridge_Schwenke_H2O/h2o_potential.py:6-7 -> calc_h2o_pot.f, h2opes_v2.f): no
Fortran compiler or `make` step is needed.  cds: (N,3,3) bohr, atoms ordered H, H, O."""
from pyvibdmc_b200 import kernels as _K, _capi


def water_pot(cds):
    return _K.pes_h2o(cds)


water_pot._pvd_builtin = {"potential": _capi.POT_H2O_PS}
