"""The named benchmark systems of BASELINE.json: component definitions and state points.

  argon  configs[0]: the component and state point of the reference's examples/Argon/200K_18mol_l (config.xml, components.xml,
         Argon_200K_18mol_l.inp: sigma = 6.4160007 a0, eps = 3.69852537e-4 Eh, m = 0.039948, T = 6.33363365e-4, rho = 1.6063e-3
         a0^-3, dt = 0.0667516), lattice generated at N = 2 * 37^3 = 101 306 molecules with rc = 2.5 sigma = 16.04 a0 (24^3 cells;
         the shipped example itself holds 2 048 molecules with rc = 5.15 sigma -- SURVEY.md 8d-1b)
  lj     configs[1]/[4]: reduced-unit 1CLJ liquid, sigma = eps = m = 1, rho* = 0.7, T* = 0.85, rc = 2.5 sigma, dt = 0.002
  water  configs[2]: the water-like model of the reference's test_input/VectorizationWater.inp:5-10 (1 LJ centre + 3 point
         charges, principal moments given), rho = 4.935e-3 a0^-3 (the fixture's 250 / 37^3), rc = rcLJ = 14.9 a0 (2.5 sigma),
         T = 9.437e-4, dt = 0.0667516
  co2    configs[3]: 2CLJQ carbon dioxide of examples/Adsorption/CO2/components_co2_stoll.xml (2 LJ centres at z = +-1.2088 A,
         point quadrupole at the centre along z), rho = 0.0137 A^-3, rc = 14.9235 A (5 sigma, as examples/Adsorption/CO2/
         run01/config.xml:38), T = 293.44 K, dt = 0.00182367; the XML ships Ixx = Iyy = Izz = 0, which would freeze the
         rotation (Component.cpp:91-95) -- here the moments follow from the site masses (Ixx = Iyy = 64.31), so molecules rotate.
"""
from __future__ import annotations

from dataclasses import dataclass

from .components import Charge, Component, Ensemble, LJcenter, Quadrupole, lj_fluid


@dataclass
class Workload:
    name: str
    ensemble: Ensemble
    rho: float
    temperature: float
    rc: float
    dt: float
    n_per_dim: int          # BCC cells per dimension of the single-GPU size: N = 2 n^3
    description: str


def water_like() -> Ensemble:
    return Ensemble([Component(
        ljcenters=[LJcenter((0.0, 0.123891518, 0.0), 0.016, 0.000246810271, 5.95953717, 0.0, False)],
        charges=[Charge((0.0, -0.159567514, 0.0), 0.0, -1.04), Charge((1.43042933, -0.983266012, 0.0), 0.001008, 0.52),
                 Charge((-1.43042933, -0.983266012, 0.0), 0.001008, 0.52)],
        I_override=(0.00219467882, 0.00412499417, 0.00631967299))])


def co2_2cljq() -> Ensemble:
    return Ensemble([Component(
        ljcenters=[LJcenter((0.0, 0.0, -1.2088), 22.005, 133.22, 2.9847, 0.0, False), LJcenter((0.0, 0.0, 1.2088), 22.005, 133.22, 2.9847, 0.0, False)],
        quadrupoles=[Quadrupole((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), -322.873584174177)])])


def argon() -> Ensemble:
    return Ensemble([Component(ljcenters=[LJcenter((0.0, 0.0, 0.0), 0.039948, 0.000369852537, 6.4160007, 16.04000175, False)])])


def get(name: str) -> Workload:
    if name == "argon":
        return Workload("argon", argon(), 0.00160630092, 0.000633363365, 16.04000175, 0.0667516, 37,
                        "BASELINE.json configs[0]: examples/Argon 1CLJ (200K_18mol_l component and state point), 101 306 molecules on a lattice, rc=2.5 sigma, mixed precision, NVT")
    if name == "lj":
        return Workload("lj", lj_fluid(rc=2.5), 0.7, 0.85, 2.5, 0.002, 200,
                        "BASELINE.json configs[1]: 1CLJ LJ liquid rho*=0.7 T*=0.85 rc=2.5 sigma dt=0.002, mixed precision, NVT velocity scaling")
    if name == "water":
        return Workload("water", water_like(), 4.935e-3, 9.437e-4, 14.9, 0.0667516, 126,
                        "BASELINE.json configs[2]: water-like model (1 LJ + 3 charges, VectorizationWater.inp), 4 M molecules, rc=2.5 sigma, mixed precision, NVT")
    if name == "co2":
        return Workload("co2", co2_2cljq(), 0.0137, 293.44, 14.9235, 0.00182367, 159,
                        "BASELINE.json configs[3]: CO2 2CLJQ (2 LJ + quadrupole, components_co2_stoll.xml), 8 M molecules, rc=5 sigma, mixed precision, NVT")
    raise ValueError("unknown workload " + name)
