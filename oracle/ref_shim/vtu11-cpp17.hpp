// Stand-in for vtu11 (VTU writer used by src/View/): nothing of it is needed to compile the solver-side headers.  TEST INFRASTRUCTURE ONLY.
#ifndef REF_SHIM_VTU11_HPP_
#define REF_SHIM_VTU11_HPP_
#include <cstdint>
#include <string>
#include <tuple>
#include <vector>
namespace vtu11 {
using VtkCellType = std::int8_t;
using VtkIndexType = std::int64_t;
using DataSetInfo = std::tuple<std::string, int, std::size_t>;
using DataSetData = std::vector<double>;
}  // namespace vtu11
#endif
