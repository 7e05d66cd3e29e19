-- ArrayTableHandler:new(size, init_value) / :get() / :add(data, sync)
-- (reference: binding/lua/ArrayTableHandler.lua:6-56; master-init protocol kept)
local ffi = require 'ffi'
local util = require('multiverso.util')
local tbh = {}
tbh.__index = tbh

function tbh:new(size, init_value)
    local mv = require('multiverso')
    local o = setmetatable({}, tbh)
    o._size = size
    o._handler = ffi.new('TableHandler[1]')
    mv.libmv.MV_NewArrayTable(size, o._handler)
    if init_value ~= nil then
        -- every worker issues a SYNC add; only the master carries the initial value
        local init = init_value
        if mv.worker_id() ~= 0 then init = torch.FloatTensor(size):zero() end
        o:add(init, true)
        mv.barrier()
    end
    return o
end

function tbh:get()
    local mv = require('multiverso')
    local cdata = ffi.new('float[?]', self._size)
    mv.libmv.MV_GetArrayTable(self._handler[0], cdata, self._size)
    return util.cdata2tensor(cdata, { self._size })
end

function tbh:add(data, sync)
    local mv = require('multiverso')
    local cdata, keep = util.tensor2cdata(data)
    if sync then
        mv.libmv.MV_AddArrayTable(self._handler[0], cdata, self._size)
    else
        mv.libmv.MV_AddAsyncArrayTable(self._handler[0], cdata, self._size)
    end
    return keep
end
return tbh
