// STdb — STAN's database file: protobuf-net 3.0.73 serialisation of the [ProtoContract] classes
// in /root/reference/src/STAN_Database.  Native reader/writer so the solver
// host can consume and produce real STAN databases without .NET.
//
//   Database   {1: map<int,Node>, 2: map<int,Element>, 3: map<int,Material>, 4: map<int,BC>,
//               5: nDOF, 6: Analysis, 7: Information}                       Database.cs:12-21
//   Node       {1: ID, 2: X, 3: Y, 4: Z, 5: EList, 6: DOF, 7-9: DispX/Y/Z}   Node.cs:11-21
//   Element    {1: ID, 2: Type, 3: PID, 4: MatID, 5: NList, 6: Strain, 7: Stress}   Element.cs:14-23
//   MatrixST   {1: M (row-major), 2: Rows, 3: Cols}                          MatrixST.cs:17-19
//   Material   {1: ID, 2: Type, 3: Name, 4: E, 5: Poisson, 6: ColorID}       Material.cs:9-14
//   BoundaryCondition {1: Type, 2: Name, 3: ID, 4: map<int,MatrixST>, 5: ColorID}   BoundaryCondition.cs:10-14
//   Analysis   {1: Type, 2: LinSolver, 3: Tolerance, 4: IterMax, 5: IncNumb, 6: Result_StepNo}   Analysis.cs:8-13
// Dictionary<int,T> = repeated {1: key, 2: value}.  Zero / empty members are omitted on the wire and
// read back as 0 / empty; repeated scalars are accepted packed or unpacked and written unpacked.
// Members this code does not interpret (Information, unknown fields) are preserved byte-for-byte.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace stdb {

struct MatrixST {
    std::vector<double> M;
    int32_t rows = 0, cols = 0;
};

struct Node {
    int32_t id = 0;
    double x = 0, y = 0, z = 0;
    std::vector<int32_t> elist, dof;
    std::vector<double> dispx, dispy, dispz;
};

struct Element {
    int32_t id = 0, pid = 0, matid = 0;
    std::string type;
    std::vector<int32_t> nlist;
    std::vector<MatrixST> strain, stress;
};

struct Material {
    int32_t id = 0, colorid = 0;
    std::string type, name;
    double E = 0, poisson = 0;
};

struct BoundaryCondition {
    std::string type, name;
    int32_t id = 0, colorid = 0;
    std::vector<std::pair<int32_t, MatrixST>> nodal;   // insertion order of the Dictionary
};

struct Analysis {
    bool present = false;
    std::string type, linsolver;
    double tolerance = 0;
    int32_t itermax = 0, incnumb = 0, result_stepno = 0;
};

struct Database {
    std::vector<Node> nodes;                  // NodeLib, insertion (= file) order
    std::vector<Element> elems;               // ElemLib
    std::vector<Material> mats;               // MatLib
    std::vector<BoundaryCondition> bcs;       // BCLib
    std::vector<int32_t> bc_keys;             // dictionary keys of BCLib (ID is stored separately)
    int32_t ndof = 0;
    Analysis analysis;
    bool has_info = false;
    std::string info_raw;                     // Information sub-message, kept verbatim
};

// Returns false and fills err on malformed input.
bool decode(const std::string &bytes, Database &db, std::string &err);
std::string encode(const Database &db);

bool read_file(const std::string &path, std::string &bytes, std::string &err);
bool write_file(const std::string &path, const std::string &bytes, std::string &err);

}  // namespace stdb
