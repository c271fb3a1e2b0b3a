"""Physical constants and unit factors (reference: classes/Constants.py:5-12, values from
scipy.constants / CODATA 2018 so that every derived table matches the reference's)."""
import scipy.constants as ct


class Constants:
    def __init__(self):
        self.hbar = ct.physical_constants['reduced Planck constant in eV s'][0] * 1e12   # eV ps / rad
        self.kb = ct.physical_constants['Boltzmann constant in eV/K'][0]                 # eV / K
        self.ev_in_J = ct.physical_constants['electron volt'][0]
        self.a_in_m = 1e-10
        self.ps_in_s = 1e-12
        self.eVpsa2_in_Wm2 = self.ev_in_J / (self.ps_in_s * (self.a_in_m) ** 2)         # eV/(ps A^2) -> W/m^2
        self.pi = ct.pi
