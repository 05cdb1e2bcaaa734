// stan_solver — native host of the linear-static path: the role of STAN_Solver.exe
// (/root/reference/src/STAN_Solver/Solver.cs:18-69 Main, :71-217 SolverLinearStatics, :454-462
// ExportOutput) on top of the C ABI of libstan_b200.so.  Reads an STdb database, numbers the DOFs,
// assembles / solves / recovers on the GPU, writes the results back into the same file format and
// prints the console lines the reference prints, so the PrePost -> Solver -> PrePost workflow is
// unchanged.  The reference's 10 s "Solver exit" sleep (Solver.cs:67-68) is not reproduced.
//
//   stan_solver model.STdb [-o out.STdb] [--strict] [--device N] [--vtu PREFIX [--vtu-ascii]]
//   stan_solver --roundtrip in.STdb out.STdb        decode + encode only (no GPU)
//   stan_solver --import-bdf mesh.bdf out.STdb      Database.ReadNastranMesh (no GPU)
//   stan_solver --build mesh.bdf out.STdb [...]     import + materials / BC text / analysis settings (no GPU)
//   stan_solver --dump in.STdb                      one-line JSON summary (no GPU)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/stan_b200.h"
#include "bdf.hpp"
#include "model_build.hpp"
#include "stdb.hpp"
#include "vtu.hpp"

namespace {

const char *SEP = "  ========================================================== ";

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int fail(const char *what) {
    fprintf(stderr, "stan_solver: %s: %s\n", what, stan_last_error());
    return 2;
}

void banner() {
    puts("");
    puts(SEP);
    puts("        STAN - STructural ANalyser : native B200 solver host");
    puts("        Solver: Linear, Statics (libstan_b200, CUDA sm_100a)");
    puts(SEP);
    puts("");
}

void summary(const stdb::Database &db) {                      // Database.Database_Summary, Database.cs:123-133
    printf("\n  ==================   DATABASE SUMMARY   ==================");
    printf("\n%-25s%31zu", "   Number of nodes:", db.nodes.size());
    printf("\n%-25s%31zu", "   Number of elements:", db.elems.size());
    printf("\n%-25s%31d", "   Number of DoF:", db.ndof);
    printf("\n%s\n", SEP);
}

int dump(const stdb::Database &db) {
    size_t with_disp = 0, with_stress = 0;
    double umax = 0, smax = 0;
    for (const auto &n : db.nodes)
        if (n.dispx.size() > 1) {
            with_disp++;
            for (const auto *v : {&n.dispx, &n.dispy, &n.dispz}) if (std::abs((*v)[1]) > umax) umax = std::abs((*v)[1]);
        }
    for (const auto &e : db.elems)
        if (e.stress.size() > 1) {
            with_stress++;
            for (double v : e.stress[1].M) if (std::abs(v) > smax) smax = std::abs(v);
        }
    printf("{\"nodes\": %zu, \"elements\": %zu, \"materials\": %zu, \"bcs\": %zu, \"ndof\": %d, \"analysis\": \"%s\", "
           "\"linsolver\": \"%s\", \"tolerance\": %.17g, \"itermax\": %d, \"result_stepno\": %d, \"nodes_with_results\": %zu, "
           "\"elements_with_results\": %zu, \"max_abs_disp\": %.17g, \"max_abs_stress\": %.17g}\n",
           db.nodes.size(), db.elems.size(), db.mats.size(), db.bcs.size(), db.ndof, db.analysis.type.c_str(),
           db.analysis.linsolver.c_str(), db.analysis.tolerance, db.analysis.itermax, db.analysis.result_stepno, with_disp,
           with_stress, umax, smax);
    return 0;
}

// Solver.SolverLinearStatics through the C ABI.  Returns 0 on success.
int linear_statics(stdb::Database &db, int device, bool strict, const std::string &vtu_prefix, bool vtu_ascii) {
    const double t_start = now_s();
    const int64_t nn = (int64_t)db.nodes.size(), ne = (int64_t)db.elems.size();
    std::unordered_map<int32_t, int32_t> node_pos, mat_pos;
    node_pos.reserve((size_t)nn * 2);
    for (int64_t i = 0; i < nn; i++) node_pos[db.nodes[i].id] = (int32_t)i;
    for (size_t i = 0; i < db.mats.size(); i++) mat_pos[db.mats[i].id] = (int32_t)i;

    std::vector<double> xyz(3 * nn);
    for (int64_t i = 0; i < nn; i++) { xyz[3 * i] = db.nodes[i].x; xyz[3 * i + 1] = db.nodes[i].y; xyz[3 * i + 2] = db.nodes[i].z; }
    std::vector<int32_t> conn(8 * ne), emat(ne);
    std::vector<uint8_t> etype(ne);
    for (int64_t e = 0; e < ne; e++) {
        const stdb::Element &el = db.elems[e];
        if (el.type == "HEX8_G2") etype[e] = STAN_HEX8_G2;
        else if (el.type == "HEX8_G1") etype[e] = STAN_HEX8_G1;
        else { fprintf(stderr, "stan_solver: element %d has type '%s'; only HEX8_G1/HEX8_G2 are on the native path\n", el.id, el.type.c_str()); return 3; }
        if (el.nlist.size() != 8) { fprintf(stderr, "stan_solver: element %d has %zu nodes\n", el.id, el.nlist.size()); return 3; }
        for (int k = 0; k < 8; k++) {
            auto it = node_pos.find(el.nlist[k]);
            if (it == node_pos.end()) { fprintf(stderr, "stan_solver: element %d references missing node %d\n", el.id, el.nlist[k]); return 3; }
            conn[8 * e + k] = it->second;
        }
        auto mt = mat_pos.find(el.matid);                       // DB.MatLib[MatID] (KeyNotFound in the reference)
        if (mt == mat_pos.end()) { fprintf(stderr, "stan_solver: element %d uses material %d which is not in MatLib\n", el.id, el.matid); return 3; }
        emat[e] = mt->second;
    }
    std::vector<double> E(db.mats.size()), nu(db.mats.size());
    for (size_t i = 0; i < db.mats.size(); i++) { E[i] = db.mats[i].E; nu[i] = db.mats[i].poisson; }

    std::vector<int32_t> spc_node, load_node;
    std::vector<double> spc_val, load_val;
    for (const auto &bc : db.bcs) {
        const bool spc = bc.type == "SPC", load = bc.type == "PointLoad";
        if (!spc && !load) continue;
        for (const auto &kv : bc.nodal) {
            auto it = node_pos.find(kv.first);
            if (it == node_pos.end()) { fprintf(stderr, "stan_solver: BC '%s' references missing node %d\n", bc.name.c_str(), kv.first); return 3; }
            double v[3] = {0, 0, 0};
            for (int d = 0; d < 3 && d < (int)kv.second.M.size(); d++) v[d] = kv.second.M[d];
            if (spc) { spc_node.push_back(it->second); spc_val.insert(spc_val.end(), v, v + 3); }
            else { load_node.push_back(it->second); load_val.insert(load_val.end(), v, v + 3); }
        }
    }

    stan_options opt = {device, 0, 1, 0};
    stan_handle *h = nullptr;
    const double t_dev0 = now_s();
    if (stan_create(&opt, &h)) return fail("stan_create");
    const double t_device_start = now_s() - t_dev0;               // CUDA context + module load: a fixed cost per process
    int rc = 0;
    std::vector<int32_t> node_index(nn);
    std::vector<double> U(3 * nn), strain((size_t)48 * ne), stress((size_t)48 * ne);
    stan_assembly_stats as;
    stan_cg_report rep;
    stan_recovery_stats rs;
    stan_cg_options cg;
    stan_chol_report chol;
    memset(&cg, 0, sizeof cg);
    memset(&rep, 0, sizeof rep);
    memset(&chol, 0, sizeof chol);
    const bool cholesky = db.analysis.present && db.analysis.linsolver == "Cholesky";   // Solver.cs:162-163
    std::string vtu_written;
    do {
        if (stan_set_mesh(h, nn, xyz.data(), ne, conn.data(), etype.data(), emat.data())) { rc = fail("stan_set_mesh"); break; }
        if (stan_set_materials(h, (int32_t)E.size(), E.data(), nu.data())) { rc = fail("stan_set_materials"); break; }
        printf("   DoF ordering: ");                              // Solver.cs:44-47
        if (stan_assign_dof(h, node_index.data())) { rc = fail("stan_assign_dof"); break; }
        printf("           Done\n");
        if (db.ndof == 0) db.ndof = (int32_t)(3 * nn);            // Database.Set_nDOF (PrePost normally stores it)
        summary(db);
        if (stan_set_spc(h, (int64_t)spc_node.size(), spc_node.data(), spc_val.data())) { rc = fail("stan_set_spc"); break; }
        if (stan_set_loads(h, (int64_t)load_node.size(), load_node.data(), load_val.data())) { rc = fail("stan_set_loads"); break; }

        printf("\n%s\n        LINEAR STATIC ANALYSIS \n%s\n", SEP, SEP);   // Solver.cs:93-95
        printf("   K Matrix assembly: ");                         // SolverFunctions.cs:127
        fflush(stdout);
        if (stan_assemble(h, &as)) { rc = fail("stan_assemble"); break; }
        printf("          Done in %.2fs\n", as.total_ms / 1000.0);   // :177

        if (cholesky) {                                           // SolverFunctions.cs:384-441
            printf("   Linear system K*U=F:\n    - Cholesky decomposition:");
            fflush(stdout);
            if (stan_solve_cholesky(h, &chol)) { rc = fail("stan_solve_cholesky"); break; }
            printf(chol.terminationtype > 0 ? "   Done\n" : "   ERROR\n");
            printf("    - Solving:                  %s termination (type %d)\n", chol.terminationtype > 0 ? "NORMAL" : "ERROR",
                   chol.terminationtype);
            printf("    Total time to solve K*U=F:  %.2fs\n", (chol.setup_ms + chol.factor_ms + chol.solve_ms) / 1000.0);
        } else {
            printf("   Solving linear system...   ");             // :273
            fflush(stdout);
            cg.epsf = db.analysis.present ? db.analysis.tolerance : 1.0e-6;   // Analysis defaults, Analysis.cs:17-20
            cg.maxits = db.analysis.present ? db.analysis.itermax : 0;
            cg.its_before_rupdate = 10;
            cg.merit_check = strict ? 0 : 1;
            if (strict && cg.maxits == 0) cg.maxits = 100000;
            if (stan_solve_cg(h, &cg, &rep)) { rc = fail("stan_solve_cg"); break; }
            printf(rep.terminationtype == 1 || rep.terminationtype == 7 ? "  NORMAL " : "  ERROR ");   // :323-325
            printf(" (type %d) in %.2fs\n", rep.terminationtype, rep.solve_ms / 1000.0);
        }

        printf("   Stress recovery: ");                           // Solver.cs:183
        fflush(stdout);
        if (stan_recover(h, &rs)) { rc = fail("stan_recover"); break; }
        if (stan_get_displacements(h, U.data())) { rc = fail("stan_get_displacements"); break; }
        if (stan_get_strain_stress(h, strain.data(), stress.data())) { rc = fail("stan_get_strain_stress"); break; }
        if (!vtu_prefix.empty()) {                                // PrePost's Export window (ExportWindow.xaml.cs:43-108)
            std::vector<float> point((size_t)24 * nn);
            std::vector<double> disp(3 * nn);
            if (stan_postprocess(h, nullptr)) { rc = fail("stan_postprocess"); break; }
            if (stan_get_scalars(h, nullptr, point.data())) { rc = fail("stan_get_scalars"); break; }
            for (int64_t i = 0; i < nn; i++)
                for (int d = 0; d < 3; d++) disp[3 * i + d] = U[3 * (int64_t)node_index[i] + d];
            std::string path, verr;
            if (!vtu::write_increment(db, disp, point, vtu_prefix, vtu_ascii, path, verr)) {
                fprintf(stderr, "stan_solver: %s\n", verr.c_str());
                rc = 2;
                break;
            }
            vtu_written = path;
        }
        printf("            Done\n");                             // :200
    } while (false);
    stan_destroy(h);
    if (rc) return rc;

    // write-back: what AssignDOF, Initialize_*, Update_Displacement and Update_StrainStress leave in the
    // object graph (Database.cs:143-158,218-223; Node.cs:95-116,176-181; Element.cs:79-113,257-267)
    for (auto &n : db.nodes) n.elist.clear();
    for (const auto &el : db.elems)
        for (int k = 0; k < 8; k++) {
            auto &l = db.nodes[node_pos[el.nlist[k]]].elist;
            if (l.empty() || l.back() != el.id) l.push_back(el.id);
        }
    for (int64_t i = 0; i < nn; i++) {
        stdb::Node &n = db.nodes[i];
        const int32_t d = 3 * node_index[i];
        n.dof = {d, d + 1, d + 2};
        n.dispx = {0.0, U[d]}; n.dispy = {0.0, U[d + 1]}; n.dispz = {0.0, U[d + 2]};
    }
    for (int64_t e = 0; e < ne; e++) {
        stdb::MatrixST zero, eps, sig;
        zero.M.assign(48, 0.0); zero.rows = 8; zero.cols = 6;
        eps.rows = sig.rows = 8; eps.cols = sig.cols = 6;
        eps.M.assign(strain.begin() + 48 * e, strain.begin() + 48 * (e + 1));
        sig.M.assign(stress.begin() + 48 * e, stress.begin() + 48 * (e + 1));
        db.elems[e].strain = {zero, eps};
        db.elems[e].stress = {zero, sig};
    }
    db.analysis.present = true;
    db.analysis.result_stepno = 1;                                // Solver.cs:56
    printf("\n%s\n  Total CPU time: %.2f s\n%s\n", SEP, now_s() - t_start, SEP);   // Solver.cs:213-216
    if (!vtu_written.empty()) printf("   Result file exported: %s\n", vtu_written.c_str());
    printf("   Device start-up (CUDA context, once per process): %.2f s\n", t_device_start);
    if (cholesky)
        printf("   Cholesky: skyline %.2f GB in %lld blocks of 64x64, factor %.2f ms (%.2f TFLOP/s), solves %.2f ms, "
               "assembly kernel %.2f ms, recovery %.2f ms\n", chol.skyline_bytes / 1e9, (long long)chol.n_blocks, chol.factor_ms,
               chol.factor_ms > 0 ? chol.flops / chol.factor_ms / 1e9 : 0.0, chol.solve_ms, as.assembly_ms, rs.recover_ms);
    else
        printf("   CG iterations: %d, ||r||/||b|| = %.3e, SpMV launches: %d, assembly kernel %.2f ms, recovery %.2f ms\n",
               rep.iterationscount, rep.bnorm > 0 ? std::sqrt(rep.r2) / rep.bnorm : 0.0, rep.spmv_launches, as.assembly_ms,
               rs.recover_ms);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    std::string err, bytes;
    if (argc >= 2 && !strcmp(argv[1], "--roundtrip")) {
        if (argc != 4) { fprintf(stderr, "usage: stan_solver --roundtrip in.STdb out.STdb\n"); return 1; }
        stdb::Database db;
        if (!stdb::read_file(argv[2], bytes, err) || !stdb::decode(bytes, db, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        if (!stdb::write_file(argv[3], stdb::encode(db), err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "--dump")) {
        if (argc != 3) { fprintf(stderr, "usage: stan_solver --dump in.STdb\n"); return 1; }
        stdb::Database db;
        if (!stdb::read_file(argv[2], bytes, err) || !stdb::decode(bytes, db, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        return dump(db);
    }
    if (argc >= 2 && !strcmp(argv[1], "--remove-results")) {      // MainWindow.RemoveResults_Click (MainWindow.xaml.cs:731-763)
        if (argc != 4) { fprintf(stderr, "usage: stan_solver --remove-results in.STdb out.STdb\n"); return 1; }
        stdb::Database db;
        if (!stdb::read_file(argv[2], bytes, err) || !stdb::decode(bytes, db, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        db.analysis.result_stepno = 0;                            // AnalysisLib.SetResultStepNo(0)
        for (auto &e : db.elems) { e.strain.clear(); e.stress.clear(); }          // Element.ClearResults: Stress = Strain = null
        for (auto &n : db.nodes) { n.dispx = {0.0}; n.dispy = {0.0}; n.dispz = {0.0}; }   // Node.Initialize_StepZero
        if (!stdb::write_file(argv[3], stdb::encode(db), err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "--build")) {               // the PrePost steps between import and solve, scripted
        if (argc < 4) {
            fprintf(stderr, "usage: stan_solver --build mesh.bdf out.STdb [--material E NU]... [--part-mat PID MATID]...\n"
                            "       [--elem-type HEX8_G1|HEX8_G2] [--spc rows.txt]... [--load rows.txt]...\n"
                            "       [--solver CG|Cholesky] [--tol T] [--itermax N]\n");
            return 1;
        }
        stdb::Database db;
        bdf::ImportReport rep;
        if (!bdf::read_nastran_mesh(argv[2], db, rep, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        db.ndof = (int32_t)(3 * db.nodes.size());
        std::string solver = "CG";
        double tol = 1.0e-6;
        int itermax = 0;
        bool part_mat_given = false;
        for (int i = 4; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--material" && i + 2 < argc) { model_build::add_material(db, atof(argv[i + 1]), atof(argv[i + 2])); i += 2; }
            else if (a == "--part-mat" && i + 2 < argc) { model_build::set_part_material(db, atoi(argv[i + 1]), atoi(argv[i + 2])); part_mat_given = true; i += 2; }
            else if (a == "--elem-type" && i + 1 < argc) { model_build::set_hex_type(db, -1, argv[++i]); }
            else if ((a == "--spc" || a == "--load") && i + 1 < argc) {
                std::string text;
                if (!stdb::read_file(argv[++i], text, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
                const bool spc = a == "--spc";
                if (!model_build::add_bc(db, spc ? "SPC" : "PointLoad", spc ? "Fix" : "Load", model_build::parse_bc_text(text), err)) {
                    fprintf(stderr, "stan_solver: %s\n", err.c_str());
                    return 3;
                }
            }
            else if (a == "--solver" && i + 1 < argc) solver = argv[++i];
            else if (a == "--tol" && i + 1 < argc) tol = atof(argv[++i]);
            else if (a == "--itermax" && i + 1 < argc) itermax = atoi(argv[++i]);
            else { fprintf(stderr, "stan_solver: unknown or incomplete option '%s'\n", a.c_str()); return 1; }
        }
        if (!part_mat_given && !db.mats.empty()) model_build::set_part_material(db, -1, db.mats.front().id);
        model_build::set_analysis(db, solver, tol, itermax);
        if (!stdb::write_file(argv[3], stdb::encode(db), err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        size_t rows = 0;
        for (const auto &bc : db.bcs) rows += bc.nodal.size();
        printf("{\"nodes\": %zu, \"elements\": %zu, \"import_errors\": %zu, \"materials\": %zu, \"bcs\": %zu, \"bc_rows\": %zu}\n",
               db.nodes.size(), db.elems.size(), rep.errors.size(), db.mats.size(), db.bcs.size(), rows);
        return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "--import-bdf")) {
        if (argc != 4) { fprintf(stderr, "usage: stan_solver --import-bdf mesh.bdf out.STdb\n"); return 1; }
        stdb::Database db;
        bdf::ImportReport rep;
        if (!bdf::read_nastran_mesh(argv[2], db, rep, err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        db.ndof = (int32_t)(3 * db.nodes.size());
        db.analysis.present = true;                               // Analysis() defaults, Analysis.cs:15-24
        db.analysis.type = "Linear_Statics"; db.analysis.linsolver = "CG"; db.analysis.tolerance = 1.0e-6;
        if (!stdb::write_file(argv[3], stdb::encode(db), err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }
        printf("{\"nodes\": %zu, \"elements\": %zu, \"import_errors\": %zu}\n", db.nodes.size(), db.elems.size(), rep.errors.size());
        return 0;
    }
    if (argc < 2) { fprintf(stderr, "usage: stan_solver model.STdb [-o out.STdb] [--strict] [--device N] [--vtu PREFIX [--vtu-ascii]]\n"); return 1; }
    std::string in = argv[1], out = argv[1];
    bool strict = false, vtu_ascii = false;
    std::string vtu_prefix;
    int device = -1;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--strict")) strict = true;
        else if (!strcmp(argv[i], "--vtu") && i + 1 < argc) vtu_prefix = argv[++i];
        else if (!strcmp(argv[i], "--vtu-ascii")) vtu_ascii = true;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    }
    banner();
    printf("   Reading input file: ");                            // Solver.cs:23-41
    stdb::Database db;
    if (!stdb::read_file(in, bytes, err) || !stdb::decode(bytes, db, err)) { fprintf(stderr, "\nstan_solver: %s\n", err.c_str()); return 2; }
    printf("     Done\n");
    if (db.analysis.present && db.analysis.type != "Linear_Statics") {
        fprintf(stderr, "stan_solver: analysis type '%s' is not on the native path (only Linear_Statics)\n", db.analysis.type.c_str());
        return 3;
    }
    if (db.analysis.present && !db.analysis.linsolver.empty() && db.analysis.linsolver != "CG" &&
        db.analysis.linsolver != "Cholesky") {                    // Solver.cs:162-164 also knows "LU"
        fprintf(stderr, "stan_solver: linear solver '%s' is not on the native path (CG and Cholesky are)\n", db.analysis.linsolver.c_str());
        return 3;
    }
    int rc = linear_statics(db, device, strict, vtu_prefix, vtu_ascii);
    if (rc) return rc;
    if (!stdb::write_file(out, stdb::encode(db), err)) { fprintf(stderr, "stan_solver: %s\n", err.c_str()); return 2; }   // ExportOutput
    return 0;
}
