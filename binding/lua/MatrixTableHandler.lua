-- MatrixTableHandler:new(num_row, num_col, init_value) / :get(row_ids) / :add(data, row_ids, sync)
-- (reference: binding/lua/MatrixTableHandler.lua:6-90)
local ffi = require 'ffi'
local util = require('multiverso.util')
local tbh = {}
tbh.__index = tbh

function tbh:new(num_row, num_col, init_value)
    local mv = require('multiverso')
    local o = setmetatable({}, tbh)
    o._num_row, o._num_col, o._size = num_row, num_col, num_row * num_col
    o._handler = ffi.new('TableHandler[1]')
    mv.libmv.MV_NewMatrixTable(num_row, num_col, o._handler)
    if init_value ~= nil then
        local init = init_value
        if mv.worker_id() ~= 0 then init = torch.FloatTensor(num_row, num_col):zero() end
        o:add(init, nil, true)
        mv.barrier()
    end
    return o
end

function tbh:get(row_ids)
    local mv = require('multiverso')
    if row_ids == nil then
        local cdata = ffi.new('float[?]', self._size)
        mv.libmv.MV_GetMatrixTableAll(self._handler[0], cdata, self._size)
        return util.cdata2tensor(cdata, { self._num_row, self._num_col })
    end
    local n = #row_ids
    local cdata = ffi.new('float[?]', n * self._num_col)
    local ids = util.tensor2cdata(row_ids, 'int')
    mv.libmv.MV_GetMatrixTableByRows(self._handler[0], cdata, n * self._num_col, ids, n)
    return util.cdata2tensor(cdata, { n, self._num_col })
end

function tbh:add(data, row_ids, sync)
    local mv = require('multiverso')
    local cdata, keep = util.tensor2cdata(data)
    if row_ids == nil then
        if sync then mv.libmv.MV_AddMatrixTableAll(self._handler[0], cdata, self._size)
        else mv.libmv.MV_AddAsyncMatrixTableAll(self._handler[0], cdata, self._size) end
    else
        local n = #row_ids
        local ids = util.tensor2cdata(row_ids, 'int')
        if sync then mv.libmv.MV_AddMatrixTableByRows(self._handler[0], cdata, n * self._num_col, ids, n)
        else mv.libmv.MV_AddAsyncMatrixTableByRows(self._handler[0], cdata, n * self._num_col, ids, n) end
    end
    return keep
end
return tbh
