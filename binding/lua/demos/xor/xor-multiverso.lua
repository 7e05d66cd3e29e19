-- XOR with a 2-20-1 MLP trained by several workers through multiverso (counterpart of the
-- reference's binding/lua/demos/xor/xor-multiverso.lua:1-88). Every worker trains on its own
-- random XOR batches and, after each step, pushes (new - old) parameters and pulls the merged
-- model: the ASGD pattern of the torch binding.
--
--   th xor-multiverso.lua                                     (1 worker)
--   python ../../../../tools/mvrun.py -n 4 -- th xor-multiverso.lua
require 'torch'
require 'nn'
local mv = require 'multiverso'

mv.init()
torch.manualSeed(1234 + mv.worker_id())

local model = nn.Sequential()
model:add(nn.Linear(2, 20)):add(nn.Tanh()):add(nn.Linear(20, 1))
local criterion = nn.MSECriterion()
local params, grads = model:getParameters()

-- one ArrayTable holds the flattened model; the master's initial values become the shared start
local table = mv.ArrayTableHandler:new(params:size(1), params:float())
params:copy(table:get())
local last = params:clone()

local batch, lr = 128, 0.01
for step = 1, 2000 do
    local x = torch.randn(batch, 2)
    local y = torch.Tensor(batch, 1)
    for i = 1, batch do y[i][1] = (x[i][1] * x[i][2] > 0) and -1 or 1 end
    grads:zero()
    local out = model:forward(x)
    local loss = criterion:forward(out, y)
    model:backward(x, criterion:backward(out, y))
    params:add(-lr, grads)
    -- sync: push the local progress, pull the global model
    table:add((params - last):float())
    params:copy(table:get())
    last:copy(params)
    if step % 200 == 0 and mv.worker_id() == 0 then
        print(string.format('step %d  loss %.4f', step, loss))
    end
end
mv.barrier()
mv.shutdown()
