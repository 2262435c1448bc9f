"""Writes ls1-MarDyn XML input (the format Simulation::readXML / Ensemble::readXML / Component::readXML parse:
src/Simulation.cpp:160-640, src/ensemble/EnsembleBase.cpp:20-90, src/molecules/Component.cpp:27-100,
src/molecules/Site.h:53-64,126-135,195-200,318-336,383-401) for an `Ensemble` of this package, so that the SAME system can
be given to the stock reference binary, to the drop-in build (adapter/) and to the C-ABI library.

All numbers are written with repr() (round-trip exact through strtod); everything is in reduced units (no unit attributes).
"""
from __future__ import annotations

from .components import Ensemble


def _site_coords(r):
    return "<coords><x>%r</x><y>%r</y><z>%r</z></coords>" % (float(r[0]), float(r[1]), float(r[2]))


def components_xml(ens: Ensemble) -> str:
    out = ["<components>"]
    for i, c in enumerate(ens.components):
        out.append('<moleculetype id="%d" name="c%d">' % (i + 1, i + 1))
        sid = 1
        for s in c.ljcenters:
            out.append('<site type="LJ126" id="%d" name="LJ%d">%s<mass>%r</mass><sigma>%r</sigma><epsilon>%r</epsilon><shifted>%s</shifted></site>'
                       % (sid, sid, _site_coords(s.r), float(s.m), float(s.sigma), float(s.eps), "true" if s.do_shift else "false"))
            sid += 1
        for s in c.charges:
            out.append('<site type="Charge" id="%d" name="q%d">%s<mass>%r</mass><charge>%r</charge></site>'
                       % (sid, sid, _site_coords(s.r), float(s.m), float(s.q)))
            sid += 1
        for s in c.dipoles:
            out.append('<site type="Dipole" id="%d" name="d%d">%s<mass>0.0</mass><dipolemoment><x>%r</x><y>%r</y><z>%r</z><abs>%r</abs></dipolemoment></site>'
                       % (sid, sid, _site_coords(s.r), float(s.e[0]), float(s.e[1]), float(s.e[2]), float(s.abs_my)))
            sid += 1
        for s in c.quadrupoles:
            out.append('<site type="Quadrupole" id="%d" name="Q%d">%s<mass>0.0</mass><quadrupolemoment><x>%r</x><y>%r</y><z>%r</z><abs>%r</abs></quadrupolemoment></site>'
                       % (sid, sid, _site_coords(s.r), float(s.e[0]), float(s.e[1]), float(s.e[2]), float(s.abs_q)))
            sid += 1
        Io = [float(x) for x in c.I_override]
        if any(x > 0.0 for x in Io):
            # Component::readXML only applies moments > 0 (Component.cpp:91-95), like the ASCII header (ASCIIReader.cpp:214-218)
            out.append('<momentsofinertia rotaxes="xyz"><Ixx>%r</Ixx><Iyy>%r</Iyy><Izz>%r</Izz></momentsofinertia>' % tuple(Io))
        out.append("</moleculetype>")
    nc = len(ens.components)
    if nc > 1:
        out.append("<mixing>")
        for i in range(nc):
            for j in range(i + 1, nc):
                out.append('<rule type="LB" cid1="%d" cid2="%d"><eta>%r</eta><xi>%r</xi></rule>' % (i + 1, j + 1, float(ens.eta[i, j]), float(ens.xi[i, j])))
        out.append("</mixing>")
    out.append("</components>")
    return "\n".join(out)


def simulation_xml(ens: Ensemble, length, temperature, dt, steps, rc, rc_lj=None, epsilon_rf=1e10, phasespace_file=None,
                   generator_density=None, datastructure="LinkedCells", integrator="Leapfrog", longrange="none",
                   result_prefix="res", extra_output="", datastructure_extra="", cells_in_cutoff=1) -> str:
    """One <mardyn> document: NVT (velocity scaling to `temperature`), periodic box, C08 LinkedCells (or the B200 container),
    ReactionField epsilon, ResultWriter with 15 digits every step. `longrange="none"` (NoLRC) keeps U_pot and p the raw
    sums of the force path; pass "homogeneous" for the reference's default."""
    rc_lj = rc if rc_lj is None else rc_lj
    if phasespace_file is not None:
        psp = '<phasespacepoint><file type="ASCII">%s</file></phasespacepoint>' % phasespace_file
    else:
        psp = ('<phasespacepoint><generator name="CubicGridGenerator"><specification>density</specification><density>%r</density>'
               "<binaryMixture>false</binaryMixture></generator></phasespacepoint>" % float(generator_density))
    out = ('<outputplugin name="ResultWriter"><writefrequency>1</writefrequency><outputprefix>%s</outputprefix>'
           "<writeprecision>15</writeprecision></outputplugin>" % result_prefix) if result_prefix else ""
    return """<?xml version='1.0' encoding='UTF-8'?>
<mardyn version="20100525">
 <refunits type="SI"><length unit="nm">0.1</length><mass unit="u">1</mass><energy unit="K">1</energy></refunits>
 <simulation type="MD">
  <integrator type="%s"><timestep unit="reduced">%r</timestep></integrator>
  <run><currenttime>0</currenttime><production><steps>%d</steps></production></run>
  <ensemble type="NVT">
   <temperature unit="reduced">%r</temperature>
   <domain type="box"><lx>%r</lx><ly>%r</ly><lz>%r</lz></domain>
   %s
   %s
  </ensemble>
  <algorithm>
   <parallelisation type="DomainDecomposition"></parallelisation>
   <datastructure type="%s"><cellsInCutoffRadius>%d</cellsInCutoffRadius>%s</datastructure>
   <cutoffs type="CenterOfMass"><defaultCutoff unit="reduced">%r</defaultCutoff><radiusLJ unit="reduced">%r</radiusLJ></cutoffs>
   <electrostatic type="ReactionField"><epsilon>%r</epsilon></electrostatic>
   <longrange type="%s"></longrange>
  </algorithm>
  <output>%s%s</output>
 </simulation>
</mardyn>
""" % (integrator, float(dt), int(steps), float(temperature), float(length[0]), float(length[1]), float(length[2]),
       components_xml(ens), psp, datastructure, int(cells_in_cutoff), datastructure_extra, float(rc), float(rc_lj), float(epsilon_rf), longrange, out, extra_output)
