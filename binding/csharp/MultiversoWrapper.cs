// multiverso-b200 :: C# binding (counterpart of the reference's C++/CLI MultiversoCLR:
// binding/C#/MultiversoCLR/MultiversoCLR.h:12-45, MultiversoCLR.cpp:23-114, MatrixTable.h).
// The reference wraps the C++ API with C++/CLI (Windows only); this is plain P/Invoke over the
// C ABI of libmultiverso.so, so it runs on .NET (Core) on Linux.  Same static surface:
// NetBind / NetConnect / NetFinalize / Init / Shutdown / CreateTable(s) / Rank / Size / Barrier /
// Get<T> / Add<T> (whole table and by row) for Int / Float / Double element types.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;

namespace MultiversoCLR
{
    public static class MultiversoWrapper
    {
        const string Lib = "multiverso";

        [DllImport(Lib)] static extern int MV_NetBindC(int rank, string endpoint);
        [DllImport(Lib)] static extern int MV_NetConnectC(int[] ranks, string[] endpoints, int size);
        [DllImport(Lib)] static extern void MV_NetFinalizeC();
        [DllImport(Lib)] static extern void MV_Init(IntPtr argc, IntPtr argv);
        [DllImport(Lib)] static extern void MV_ShutDownEx(int finalizeNet);
        [DllImport(Lib)] static extern void MV_Barrier();
        [DllImport(Lib)] static extern int MV_Rank();
        [DllImport(Lib)] static extern int MV_Size();
        [DllImport(Lib)] static extern int MV_NumWorkers();
        [DllImport(Lib)] static extern int MV_WorkerId();
        [DllImport(Lib)] static extern int MV_ServerId();
        [DllImport(Lib)] static extern int MV_SetFlagBool(string name, int value);
        [DllImport(Lib)] static extern void MV_NewMatrixTable64(long numRow, long numCol, int dtype, int isSparse,
            int isPipeline, int randomInit, double minValue, double maxValue, out IntPtr handle);
        [DllImport(Lib)] static extern void MV_GetMatrixTable64(IntPtr h, int dtype, IntPtr data, long size,
            long[] rowIds, long rowIdsN, int workerIdOpt);
        [DllImport(Lib)] static extern void MV_AddMatrixTable64(IntPtr h, int dtype, IntPtr data, long size,
            long[] rowIds, long rowIdsN, IntPtr addOption20, int isAsync);

        struct Table { public IntPtr Handle; public int Dtype; public long Rows, Cols; }
        static readonly List<Table> tables = new List<Table>();

        static int DtypeOf(string eleType)
        {
            switch (eleType) { case "Float": return 0; case "Double": return 1; case "Int": return 2; }
            throw new ArgumentException("element type must be Int, Float or Double");
        }
        static int DtypeOf<T>()
        {
            if (typeof(T) == typeof(float)) return 0;
            if (typeof(T) == typeof(double)) return 1;
            if (typeof(T) == typeof(int)) return 2;
            throw new ArgumentException("unsupported element type");
        }

        public static int NetBind(int rank, string endpoint) { return MV_NetBindC(rank, endpoint); }
        public static int NetConnect(int[] ranks, string[] endpoints) { return MV_NetConnectC(ranks, endpoints, ranks.Length); }
        public static void NetFinalize() { MV_NetFinalizeC(); }

        public static void Init(int numTables, bool sync)
        {
            MV_SetFlagBool("sync", sync ? 1 : 0);
            MV_Init(IntPtr.Zero, IntPtr.Zero);
            tables.Capacity = Math.Max(tables.Capacity, numTables);
        }
        // the reference calls MV_ShutDown(false): the net stays up so Init can be called again
        public static void Shutdown() { MV_ShutDownEx(0); tables.Clear(); }
        public static int Rank() { return MV_Rank(); }
        public static int Size() { return MV_Size(); }
        public static void Barrier() { MV_Barrier(); }

        public static void CreateTables(int[] rows, int[] cols, string[] eleTypes)
        {
            for (int i = 0; i < rows.Length; ++i) CreateTable(i, rows[i], cols[i], eleTypes[i]);
        }
        public static void CreateTable(int tableId, int rows, int cols, string eleType)
        {
            var t = new Table { Dtype = DtypeOf(eleType), Rows = rows, Cols = cols };
            MV_NewMatrixTable64(rows, cols, t.Dtype, 0, 0, 0, 0.0, 0.0, out t.Handle);
            while (tables.Count <= tableId) tables.Add(new Table());
            tables[tableId] = t;
        }

        public static void Get<T>(int tableId, T[] value) where T : struct
        {
            var t = tables[tableId];
            var pin = GCHandle.Alloc(value, GCHandleType.Pinned);
            try { MV_GetMatrixTable64(t.Handle, DtypeOf<T>(), pin.AddrOfPinnedObject(), value.Length, null, 0, -1); }
            finally { pin.Free(); }
        }
        public static void Get<T>(int tableId, int rowId, T[] value) where T : struct
        {
            var t = tables[tableId];
            var pin = GCHandle.Alloc(value, GCHandleType.Pinned);
            try { MV_GetMatrixTable64(t.Handle, DtypeOf<T>(), pin.AddrOfPinnedObject(), value.Length, new long[] { rowId }, 1, -1); }
            finally { pin.Free(); }
        }
        public static void Add<T>(int tableId, T[] update) where T : struct
        {
            var t = tables[tableId];
            var pin = GCHandle.Alloc(update, GCHandleType.Pinned);
            try { MV_AddMatrixTable64(t.Handle, DtypeOf<T>(), pin.AddrOfPinnedObject(), update.Length, null, 0, IntPtr.Zero, 0); }
            finally { pin.Free(); }
        }
        public static void Add<T>(int tableId, int rowId, T[] update) where T : struct
        {
            var t = tables[tableId];
            var pin = GCHandle.Alloc(update, GCHandleType.Pinned);
            try { MV_AddMatrixTable64(t.Handle, DtypeOf<T>(), pin.AddrOfPinnedObject(), update.Length, new long[] { rowId }, 1, IntPtr.Zero, 0); }
            finally { pin.Free(); }
        }
    }
}
