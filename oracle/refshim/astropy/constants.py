"""`const.X.cgs.value` / `const.X.esu.value` for the constants source/phys_const.py reads.  Values: astropy 4.x CODATA 2014
set in cgs, the same numbers helios_b200/host.py carries, so that both sides of a comparison evaluate with identical
scalars (kernels use their own literals, K:36-41)."""
import types


def _c(v):
    return types.SimpleNamespace(cgs=types.SimpleNamespace(value=v), esu=types.SimpleNamespace(value=v), value=v)


c = _c(29979245800.0)
k_B = _c(1.38064852e-16)
h = _c(6.62607004e-27)
R = _c(83144598.0)
N_A = _c(6.022140857e23)
sigma_sb = _c(5.6703669999999995e-05)
au = _c(14959787070000.0)
u = _c(1.6605390666e-24)
m_e = _c(9.10938356e-28)
e = _c(4.80320425e-10)
R_sun = _c(69570000000.0)
M_sun = _c(1.9884754153381438e33)
R_jup = _c(7149200000.0)
M_jup = _c(1.8981871658715508e30)
R_earth = _c(637810000.0)
M_earth = _c(5.972364730419773e27)
G = _c(6.67408e-08)
sigma_T = _c(6.6524587158e-25)
