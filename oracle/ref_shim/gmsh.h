// stand-in for the Gmsh SDK header: the entry points the reference's Mesh headers name.  The golden-vector driver never builds a
// mesh, so none of them is ever called; they only have to parse.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <utility>
#include <vector>
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
template <typename... A> inline void getNodes(A&&...) {}
template <typename... A> inline void getElements(A&&...) {}
template <typename... A> inline void createEdges(A&&...) {}
template <typename... A> inline void createFaces(A&&...) {}
template <typename... A> inline void getBasisFunctions(A&&...) {}
template <typename... A> inline void getElementProperties(A&&...) {}
template <typename... A> inline void getIntegrationPoints(A&&...) {}
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
