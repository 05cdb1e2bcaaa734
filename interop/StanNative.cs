// StanNative.cs — P/Invoke binding of libstan_b200.so / stan_b200.dll (include/stan_b200.h) and the
// replacement body of Solver.SolverLinearStatics (src/STAN_Solver/Solver.cs:97-210).
// SOURCE ONLY: the build image has no .NET toolchain, so this file is not compiled or tested here.
// Marshalling: every array is a blittable double[]/int[]/byte[] pinned for the duration of one
// call; the native side copies before returning and keeps no managed pointer.
using System;
using System.Collections.Generic;
using System.Linq;
using System.Runtime.InteropServices;
using STAN_Database;

namespace STAN_Solver
{
    internal static class StanNative
    {
        private const string Lib = "stan_b200";   // libstan_b200.so on Linux, stan_b200.dll on Windows

        [StructLayout(LayoutKind.Sequential)]
        internal struct Options { public int device, rank, world, flags; }

        [StructLayout(LayoutKind.Sequential)]
        internal struct CgOptions
        {
            public double epsf; public int maxits, its_before_rupdate, its_before_restart, merit_check,
                zero_based_counter, time_kernels, reserved;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct CgReport
        {
            public int terminationtype, iterationscount, nmv, spmv_launches;
            public double r2, bnorm, solve_ms, spmv_ms;
            public long spmv_bytes, iter_bytes, kernel_launches;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct AssemblyStats
        {
            public long n_dof, n_fixed, n_rows_local, n_blocks_local, nnz_upper, assembly_bytes;
            public double assembly_flops, pattern_ms, assembly_ms, total_ms;
            public long kernel_launches;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct CholReport
        {
            public int terminationtype, block;
            public long n, n_blocks, skyline_bytes;
            public double flops, setup_ms, factor_ms, solve_ms;
            public long kernel_launches;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct RecoveryStats { public double recover_ms; public long recover_bytes, kernel_launches; }

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern IntPtr stan_last_error();
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_create(ref Options o, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_destroy(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_set_mesh(IntPtr h, long nNodes, double[] xyz, long nElem, int[] conn, byte[] elemType, int[] elemMat);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_set_materials(IntPtr h, int nMat, double[] E, double[] nu);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_set_dof_map(IntPtr h, int[] nodeIndex);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_set_spc(IntPtr h, long n, int[] node, double[] val3);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_set_loads(IntPtr h, long n, int[] node, double[] fxyz);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_assemble(IntPtr h, out AssemblyStats st);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_solve_cg(IntPtr h, ref CgOptions o, out CgReport rep);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_solve_cholesky(IntPtr h, out CholReport rep);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_recover(IntPtr h, out RecoveryStats st);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_get_displacements(IntPtr h, double[] uFull);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_get_node_displacements(IntPtr h, double[] disp);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] internal static extern int stan_get_strain_stress(IntPtr h, double[] strain, double[] stress);

        internal static void Check(int rc)
        {
            if (rc != 0) throw new InvalidOperationException("libstan_b200 error " + rc + ": " + Marshal.PtrToStringAnsi(stan_last_error()));
        }
    }

    // Drop-in for the numeric body of Solver.SolverLinearStatics.  Call after the Initialize_* loops
    // (Solver.cs:81-90); it leaves Node.dU_buffer and Element.dE/dS filled exactly as the managed
    // code at Solver.cs:171-196 would, so Update_Displacement / Update_StrainStress (:203-210) and
    // ExportOutput run unchanged.
    internal static class StanNativeLinearStatics
    {
        internal static void Run(Database DB)
        {
            var sw = System.Diagnostics.Stopwatch.StartNew();
            // ---- flatten the object graph in dictionary (= file) order ----
            var nodes = DB.NodeLib.Values.ToList();
            var elems = DB.ElemLib.Values.ToList();
            var nodePos = new Dictionary<int, int>(nodes.Count);
            for (int i = 0; i < nodes.Count; i++) nodePos[nodes[i].ID] = i;
            var matIds = DB.MatLib.Keys.ToList();
            var matPos = new Dictionary<int, int>();
            for (int i = 0; i < matIds.Count; i++) matPos[matIds[i]] = i;

            var xyz = new double[3 * nodes.Count];
            var nodeIndex = new int[nodes.Count];
            for (int i = 0; i < nodes.Count; i++)
            {
                xyz[3 * i] = nodes[i].X; xyz[3 * i + 1] = nodes[i].Y; xyz[3 * i + 2] = nodes[i].Z;
                nodeIndex[i] = nodes[i].DOF[0] / 3;                        // Node.SetDOF, Node.cs:218-223
            }
            var conn = new int[8 * elems.Count];
            var etype = new byte[elems.Count];
            var emat = new int[elems.Count];
            for (int e = 0; e < elems.Count; e++)
            {
                for (int k = 0; k < 8; k++) conn[8 * e + k] = nodePos[elems[e].NList[k]];
                etype[e] = (byte)(elems[e].Type == "HEX8_G1" ? 1 : 2);
                emat[e] = matPos[elems[e].MatID];
            }
            var E = matIds.Select(id => DB.MatLib[id].E).ToArray();
            var nu = matIds.Select(id => DB.MatLib[id].Poisson).ToArray();

            var spcNode = new List<int>(); var spcVal = new List<double>();
            foreach (var BC in DB.BCLib.Values.Where(x => x.Type == "SPC"))
                foreach (var kv in BC.NodalValues)
                { spcNode.Add(nodePos[kv.Key]); for (int d = 0; d < 3; d++) spcVal.Add(kv.Value.Get(d, 0)); }
            var loadNode = new List<int>(); var loadVal = new List<double>();
            foreach (var BC in DB.BCLib.Values.Where(x => x.Type == "PointLoad"))
                foreach (var kv in BC.NodalValues)
                { loadNode.Add(nodePos[kv.Key]); for (int d = 0; d < 3; d++) loadVal.Add(kv.Value.Get(d, 0)); }

            var opts = new StanNative.Options { device = -1, rank = 0, world = 1, flags = 0 };
            StanNative.Check(StanNative.stan_create(ref opts, out IntPtr h));
            try
            {
                StanNative.Check(StanNative.stan_set_mesh(h, nodes.Count, xyz, elems.Count, conn, etype, emat));
                StanNative.Check(StanNative.stan_set_materials(h, E.Length, E, nu));
                StanNative.Check(StanNative.stan_set_dof_map(h, nodeIndex));
                StanNative.Check(StanNative.stan_set_spc(h, spcNode.Count, spcNode.ToArray(), spcVal.ToArray()));
                StanNative.Check(StanNative.stan_set_loads(h, loadNode.Count, loadNode.ToArray(), loadVal.ToArray()));

                Console.Write("   K Matrix assembly: ");                   // same console lines as SolverFunctions.cs:127,177
                StanNative.Check(StanNative.stan_assemble(h, out var a));
                Console.WriteLine("          Done in " + (a.total_ms / 1000.0).ToString("F2") + "s");

                if (DB.AnalysisLib.GetLinSolver() == "Cholesky")           // Solver.cs:163, SolverFunctions.cs:384-441
                {
                    Console.WriteLine("   Linear system K*U=F:");
                    Console.Write("    - Cholesky decomposition:");
                    StanNative.Check(StanNative.stan_solve_cholesky(h, out var ch));
                    Console.WriteLine(ch.terminationtype > 0 ? "   Done" : "   ERROR");
                    Console.WriteLine("    - Solving:                  " + (ch.terminationtype > 0 ? "NORMAL" : "ERROR") +
                                      " termination (type " + ch.terminationtype + ")");
                    Console.WriteLine("    Total time to solve K*U=F:  " + ((ch.setup_ms + ch.factor_ms + ch.solve_ms) / 1000.0).ToString("F2") + "s");
                }
                else
                {
                    Console.Write("   Solving linear system...   ");            // SolverFunctions.cs:273
                    var cg = new StanNative.CgOptions
                    {
                        epsf = DB.AnalysisLib.GetLinSolverTolerance(), maxits = DB.AnalysisLib.GetLinSolverMaxIter(),
                        its_before_rupdate = 10, its_before_restart = 0, merit_check = 1
                    };
                    StanNative.Check(StanNative.stan_solve_cg(h, ref cg, out var rep));
                    Console.Write(rep.terminationtype == 1 || rep.terminationtype == 7 ? "  NORMAL " : "  ERROR ");   // :323-325
                    Console.WriteLine(" (type " + rep.terminationtype + ") in " + (rep.solve_ms / 1000.0).ToString("F2") + "s");
                }

                Console.Write("   Stress recovery: ");                     // Solver.cs:183
                StanNative.Check(StanNative.stan_recover(h, out _));
                var disp = new double[3 * nodes.Count];                    // Node.dU_buffer per node, NodeLib order
                var strain = new double[48 * elems.Count];
                var stress = new double[48 * elems.Count];
                StanNative.Check(StanNative.stan_get_node_displacements(h, disp));
                StanNative.Check(StanNative.stan_get_strain_stress(h, strain, stress));
                Console.WriteLine("            Done");

                for (int i = 0; i < nodes.Count; i++)                      // Solver.cs:171-178 (the gather ran on the device)
                    for (int d = 0; d < 3; d++) nodes[i].dU_buffer[d] = disp[3 * i + d];
                for (int e = 0; e < elems.Count; e++)                      // what Recovery_Stress leaves in dE/dS
                    for (int i = 0; i < 8; i++)
                        for (int c = 0; c < 6; c++)
                        {
                            elems[e].dE[i].SetFast(c, 0, strain[48 * e + 6 * i + c]);
                            elems[e].dS[i].SetFast(c, 0, stress[48 * e + 6 * i + c]);
                        }
            }
            finally { StanNative.stan_destroy(h); }
            sw.Stop();
        }
    }
}
