"""Writes a g++-compilable copy of ONE reference header into a scratch directory (never into the repository).

src/Solver/VariableConvertor.cpp:228-231 specialises a member template inside its class (`template <> getScalar<VelocitySquaredNorm>`),
which icpx / clang accept and g++ rejects (CWG 727).  The copy folds that specialisation into the primary template with `if constexpr`
— same arithmetic, same call sites.  usage: python ref_patch.py /root/reference/src OUTDIR"""
import os
import sys

src, out = sys.argv[1], sys.argv[2]
path = os.path.join(src, "Solver", "VariableConvertor.cpp")
text = open(path).read()
primary = """  template <ComputationalVariableEnum ComputationalVariableType>
  [[nodiscard]] inline Real getScalar(const Isize column) const {
    return this->computational_(getComputationalVariableIndex<SimulationControl, ComputationalVariableType>(), column);
  }
"""
special = """  template <>
  [[nodiscard]] inline Real getScalar<ComputationalVariableEnum::VelocitySquaredNorm>(const Isize column) const {
    return this->getVector<ComputationalVariableEnum::Velocity>(column).squaredNorm();
  }
"""
if text.count(primary) != 1 or text.count(special) != 1:
    raise SystemExit("reference source does not look as expected: VariableConvertor.cpp getScalar<ComputationalVariableEnum>")
folded = """  template <ComputationalVariableEnum ComputationalVariableType>
  [[nodiscard]] inline Real getScalar(const Isize column) const {
    if constexpr (ComputationalVariableType == ComputationalVariableEnum::VelocitySquaredNorm) {
      return this->template getVector<ComputationalVariableEnum::Velocity>(column).squaredNorm();
    } else {
      return this->computational_(getComputationalVariableIndex<SimulationControl, ComputationalVariableType>(), column);
    }
  }
"""
text = text.replace(special, "").replace(primary, folded)
os.makedirs(os.path.join(out, "Solver"), exist_ok=True)
open(os.path.join(out, "Solver", "VariableConvertor.cpp"), "w").write(text)
