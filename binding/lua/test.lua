-- testArray / testMatrix (reference: binding/lua/test.lua:16-74); run with `luajit test.lua`
local mv = require('multiverso')
mv.init(true)
local size = 1000
local tbh = mv.ArrayTableHandler:new(size)
for i = 1, 10 do
    tbh:add(torch.range(1, size), true)
    tbh:add(torch.range(1, size), true)
    local got = tbh:get()
    assert(math.abs(got[5] - 5 * i * 2 * mv.num_workers()) < 1e-3)
end
local m = mv.MatrixTableHandler:new(11, 10)
local base = torch.range(0, 109):resize(11, 10)
m:add(base, nil, true)
m:add(base:index(1, torch.LongTensor({1, 2, 6, 11})), {0, 1, 5, 10}, true)
local g = m:get({0, 5})
assert(math.abs(g[2][3] - base[6][3] * 2 * mv.num_workers()) < 1e-3)
mv.barrier()
mv.shutdown()
print('lua binding ok')
