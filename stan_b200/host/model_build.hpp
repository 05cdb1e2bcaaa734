// What a user does in PrePost between "import mesh" and "solve" (README.md:50-76 of the reference),
// without the GUI: materials, part properties, boundary conditions pasted as text, analysis settings.
#pragma once
#include <string>
#include <vector>

#include "stdb.hpp"

namespace model_build {

// BOX_BC.Paste_Click (/root/reference/src/STAN_PrePost/BOX_BC.xaml.cs:228-270): one row per line,
// four fields "NID x y z" separated by ',' or ' ' or TAB (tried in that order, the first that yields
// exactly four fields wins), invariant-culture numbers, unparsable lines skipped silently, and — as in
// the reference — nothing is read unless the text has at least two lines.
struct BcRow { int32_t nid; double v[3]; };
std::vector<BcRow> parse_bc_text(const std::string &text);

// MainWindow.AddMat + BOX_Mat: Material(ID) { Type "Elastic", ColorID = ID % 9, Name "New Material" }.
void add_material(stdb::Database &db, double E, double poisson);
// Part.Set_MatID / Part.Assign_FEtype: properties of every element of a part (pid < 0 = all parts).
void set_part_material(stdb::Database &db, int32_t pid, int32_t matid);
void set_hex_type(stdb::Database &db, int32_t pid, const std::string &type);
// MainWindow.AddBC + BOX_BC.Apply_Click + BoundaryCondition.Add (BoundaryCondition.cs:87-98): rows whose
// node is not in NodeLib are dropped; a node listed twice is Dictionary.Add's ArgumentException.
bool add_bc(stdb::Database &db, const std::string &type, const std::string &name, const std::vector<BcRow> &rows,
            std::string &err);
// Analysis() defaults (Analysis.cs:15-24) with the three fields the GUI edits (BOX_Analysis).
void set_analysis(stdb::Database &db, const std::string &linsolver, double tolerance, int32_t itermax);

}  // namespace model_build
