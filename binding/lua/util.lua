-- tensor <-> C data helpers (reference: binding/lua/util.lua:7-32)
local ffi = require 'ffi'
local util = {}
util.tensor_type = { ['torch.FloatTensor'] = 'float*', ['torch.CudaTensor'] = 'float*' }

function util.tensor2cdata(data, data_type)
    if type(data) == 'table' then
        local arr = ffi.new((data_type or 'float') .. '[?]', #data)
        for i = 1, #data do arr[i - 1] = data[i] end
        return arr
    end
    data = data:contiguous():float()
    return ffi.cast('float*', data:data()), data
end

function util.cdata2tensor(cdata, sizes)
    local n = 1
    for _, s in ipairs(sizes) do n = n * s end
    local t = torch.FloatTensor(n)
    ffi.copy(t:data(), cdata, n * ffi.sizeof('float'))
    return t:resize(table.unpack(sizes))
end
return util
