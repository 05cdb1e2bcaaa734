// VTU export of a solved database: what PrePost's Export window writes through
// vtkXMLUnstructuredGridWriter (/root/reference/src/STAN_PrePost/ExportWindow.xaml.cs:43-108,
// Part.ExportGrid, /root/reference/src/STAN_Database/Part.cs:857-939).
#pragma once
#include <string>
#include <vector>

#include "stdb.hpp"

namespace vtu {

// point[node position in db.nodes][24]: the 24 nodal-averaged fields of Part.Load_Scalar in its
// array order (stan_get_scalars).  disp[node position][3].  Writes <prefix>_001.vtu (increment 1).
bool write_increment(const stdb::Database &db, const std::vector<double> &disp, const std::vector<float> &point,
                     const std::string &prefix, bool ascii, std::string &path_out, std::string &err);

}  // namespace vtu
