// Functional stand-in for the three Gmsh SDK calls the reference's basis-function and quadrature classes make in their constructors
// (src/Mesh/Quadrature.cpp:27-35, src/Mesh/BasisFunction.cpp:30-80): getIntegrationPoints("Gauss<n>"), getBasisFunctions("[Grad]Lagrange<g>" /
// "[Grad]H1Legendre<p>"), getElementProperties.  The answers come from oracle/tables.hpp — the repository's restatement of Gmsh 4.13.1's
// tables (pinned by exactness properties and by the reference's embedded literals) — in Gmsh's output layouts, so that the
// REFERENCE'S OWN CODE assembles its operator tables, projects the initial condition and runs its sweeps (oracle/ref_sweeps.cpp).
// Every other Gmsh entry point only has to parse (meshes are handed over as arrays).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../tables.hpp"

namespace gmsh {
typedef std::vector<std::pair<int, int>> vectorpair;
template <typename... A> inline void initialize(A&&...) {}
template <typename... A> inline void finalize(A&&...) {}
template <typename... A> inline void clear(A&&...) {}
template <typename... A> inline void open(A&&...) {}
template <typename... A> inline void write(A&&...) {}
namespace option { template <typename... A> inline void setNumber(A&&...) {} template <typename... A> inline void getNumber(A&&...) {} template <typename... A> inline void getString(A&&...) {} }
namespace model {
template <typename... A> inline void add(A&&...) {}
template <typename... A> inline void getPhysicalGroups(A&&...) {}
template <typename... A> inline void getPhysicalName(A&&...) {}
template <typename... A> inline void getEntitiesForPhysicalGroup(A&&...) {}
namespace mesh {

// Gmsh element type number -> (ElementEnum value, order): SimulationControl.cpp:26-31
inline void decodeType(int gmshType, int& type, int& order) {
  static const int line[5] = {1, 8, 26, 27, 28}, tri[5] = {2, 9, 21, 23, 25}, quad[5] = {3, 10, 36, 37, 38}, hex[5] = {5, 12, 92, 93, 94};
  if (gmshType == 15) { type = orc::kPoint; order = 0; return; }
  for (int p = 0; p < 5; p++) {
    if (gmshType == line[p]) { type = orc::kLine; order = p + 1; return; }
    if (gmshType == tri[p]) { type = orc::kTriangle; order = p + 1; return; }
    if (gmshType == quad[p]) { type = orc::kQuadrangle; order = p + 1; return; }
    if (gmshType == hex[p]) { type = orc::kHexahedron; order = p + 1; return; }
  }
  type = -1; order = 0;   // tetrahedron / pyramid: MeshData<SC, 3> constructs their (unused) table objects as well -> zero-filled answers
}

inline void getIntegrationPoints(const int elementType, const std::string& integrationType, std::vector<double>& localCoord, std::vector<double>& weights) {
  int type, order; decodeType(elementType, type, order);
  if (type < 0) { localCoord.assign(3 * 1024, 0.0); weights.assign(1024, 0.0); return; }
  if (integrationType.rfind("Gauss", 0) != 0) throw std::runtime_error("gmsh stand-in: integration type " + integrationType);
  // a rule the restatement does not tabulate (triangle faces of order 7: only tetrahedra / pyramids, which are not built, would use them)
  if (type == orc::kTriangle && std::stoi(integrationType.substr(5)) > 6) { localCoord.assign(3 * 1024, 0.0); weights.assign(1024, 0.0); return; }
  const orc::Quadrature q = orc::makeQuadrature(type, std::stoi(integrationType.substr(5)));
  localCoord = q.pts; weights = q.wts;
}

// basisFunctions: [point][function] for values, [point][function][3] for gradients (Gmsh's layout, read back at BasisFunction.cpp:125-131,203-229)
inline void getBasisFunctions(const int elementType, const std::vector<double>& localCoord, const std::string& functionSpaceType, int& numComponents,
                              std::vector<double>& basisFunctions, int& numOrientations, const std::vector<int>& = std::vector<int>()) {
  int type, order; decodeType(elementType, type, order);
  std::string name = functionSpaceType;
  const bool grad = name.rfind("Grad", 0) == 0;
  if (grad) name = name.substr(4);
  const std::size_t npt = localCoord.size() / 3;
  std::vector<double> val; std::vector<std::array<double, 3>> g;
  numComponents = grad ? 3 : 1; numOrientations = 1;
  basisFunctions.clear();
  if (type < 0) { basisFunctions.assign(npt * 256 * 3, 0.0); return; }
  auto emit = [&](auto& basis) {
    for (std::size_t i = 0; i < npt; i++) {
      basis.eval(localCoord[3 * i], localCoord[3 * i + 1], localCoord[3 * i + 2], val, g);
      for (std::size_t b = 0; b < val.size(); b++) {
        if (grad) for (int k = 0; k < 3; k++) basisFunctions.push_back(g[b][k]);
        else basisFunctions.push_back(val[b]);
      }
    }
  };
  if (name.rfind("Lagrange", 0) == 0) {
    if (type == orc::kPoint) { for (std::size_t i = 0; i < npt; i++) { if (grad) for (int k = 0; k < 3; k++) basisFunctions.push_back(0.0); else basisFunctions.push_back(1.0); } return; }
    orc::LagrangeBasis lb(type, std::stoi(name.substr(8)));
    emit(lb);
  } else if (name.rfind("H1Legendre", 0) == 0) {
    orc::ModalBasis mb(type, std::stoi(name.substr(10)));
    emit(mb);
  } else {
    throw std::runtime_error("gmsh stand-in: function space " + functionSpaceType);
  }
}

inline void getElementProperties(const int elementType, std::string& elementName, int& dim, int& order, int& numNodes, std::vector<double>& localNodeCoord,
                                 int& numPrimaryNodes) {
  int type; decodeType(elementType, type, order);
  if (type < 0) { dim = 3; numNodes = 0; numPrimaryNodes = 0; elementName = "unused"; localNodeCoord.assign(1024, 0.0); return; }
  dim = orc::elemDim(type); elementName = "element";
  if (type == orc::kPoint) { numNodes = 1; numPrimaryNodes = 1; localNodeCoord.assign(1, 0.0); return; }
  const auto nodes = orc::referenceNodes(type, order);
  numNodes = (int)nodes.size(); numPrimaryNodes = orc::numCorners(type);
  localNodeCoord.clear();
  for (const auto& x : nodes) for (int k = 0; k < dim; k++) localNodeCoord.push_back(x[k]);   // Gmsh: dim coordinates per node
}

template <typename... A> inline void getNodes(A&&...) {}
template <typename... A> inline void getElements(A&&...) {}
template <typename... A> inline void createEdges(A&&...) {}
template <typename... A> inline void createFaces(A&&...) {}
template <typename... A> inline void getJacobian(A&&...) {}
template <typename... A> inline void getJacobians(A&&...) {}
template <typename... A> inline void getElementQualities(A&&...) {}
template <typename... A> inline void getElementsByType(A&&...) {}
template <typename... A> inline void getElementEdgeNodes(A&&...) {}
template <typename... A> inline void getElementFaceNodes(A&&...) {}
template <typename... A> inline void getEdges(A&&...) {}
template <typename... A> inline void getFaces(A&&...) {}
template <typename... A> inline void getPeriodic(A&&...) {}
template <typename... A> inline void getPeriodicNodes(A&&...) {}
}  // namespace mesh
}  // namespace model
}  // namespace gmsh
