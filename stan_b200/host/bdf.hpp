// Nastran short-format import as STAN does it: Database.ReadNastranMesh
// (/root/reference/src/STAN_Database/Database.cs:39-111), Node(string) (Node.cs:25-80),
// Element(string) (Element.cs:35-73).  Quirks are kept: only CHEXA is accepted; GRID lines are cut
// into 8-character fields, blank fields are dropped, exponent-less floats ("-7.11-15") are patched
// for '-' but not for '+' (the reference discards the result of its Replace, so such nodes fail to
// parse and are skipped); element fields are split on whitespace and every '+' is removed.
#pragma once
#include <string>
#include <vector>

#include "stdb.hpp"

namespace bdf {

struct ImportReport {
    std::vector<std::string> errors;   // Database.Import_Error
};

bool read_nastran_mesh(const std::string &path, stdb::Database &db, ImportReport &rep, std::string &err);

}  // namespace bdf
