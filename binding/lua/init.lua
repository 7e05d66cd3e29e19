-- multiverso-b200 :: Lua/Torch binding (counterpart of the reference's binding/lua/init.lua:7-67).
-- LuaJIT FFI over the C API of libmultiverso.so (include/multiverso/c_api.h -- the same float
-- array / matrix entry points as the reference, so this file works against either library).
local ffi = require 'ffi'
local mv = {}

ffi.cdef[[
    typedef void* TableHandler;
    void MV_Init(int* argc, char* argv[]);
    void MV_ShutDown();
    void MV_Barrier();
    int MV_NumWorkers();
    int MV_WorkerId();
    int MV_ServerId();
    void MV_NewArrayTable(int size, TableHandler* out);
    void MV_GetArrayTable(TableHandler handler, float* data, int size);
    void MV_AddArrayTable(TableHandler handler, float* data, int size);
    void MV_AddAsyncArrayTable(TableHandler handler, float* data, int size);
    void MV_NewMatrixTable(int num_row, int num_col, TableHandler* out);
    void MV_GetMatrixTableAll(TableHandler handler, float* data, int size);
    void MV_AddMatrixTableAll(TableHandler handler, float* data, int size);
    void MV_AddAsyncMatrixTableAll(TableHandler handler, float* data, int size);
    void MV_GetMatrixTableByRows(TableHandler handler, float* data, int size, int row_ids[], int row_ids_n);
    void MV_AddMatrixTableByRows(TableHandler handler, float* data, int size, int row_ids[], int row_ids_n);
    void MV_AddAsyncMatrixTableByRows(TableHandler handler, float* data, int size, int row_ids[], int row_ids_n);
]]

-- libmultiverso.so is searched on package.cpath like the reference does
local function load_lib()
    local name = 'libmultiverso.so'
    for path in string.gmatch(package.cpath, '[^;]+') do
        local candidate = path:gsub('%?%.so', ''):gsub('%?', '') .. name
        local f = io.open(candidate, 'r')
        if f then f:close(); return ffi.load(candidate, true) end
    end
    return ffi.load('multiverso', true)
end
mv.libmv = load_lib()

mv.util = require('multiverso.util')
mv.ArrayTableHandler = require('multiverso.ArrayTableHandler')
mv.MatrixTableHandler = require('multiverso.MatrixTableHandler')

function mv.init(sync)
    -- argv = {"", "-sync=true"} exactly like the reference binding
    sync = sync or false
    local args = {''}
    if sync then args[#args + 1] = '-sync=true' end
    local argc = ffi.new('int[1]', #args)
    local argv = ffi.new('char*[?]', #args)
    for i = 1, #args do
        argv[i - 1] = ffi.new('char[?]', #args[i] + 1)
        ffi.copy(argv[i - 1], args[i])
    end
    mv.libmv.MV_Init(argc, argv)
end
function mv.barrier() mv.libmv.MV_Barrier() end
function mv.shutdown() mv.libmv.MV_ShutDown() end
function mv.num_workers() return mv.libmv.MV_NumWorkers() end
function mv.worker_id() return mv.libmv.MV_WorkerId() end
function mv.server_id() return mv.libmv.MV_ServerId() end

return mv
