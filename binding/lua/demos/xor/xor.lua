-- Single-process baseline of xor-multiverso.lua (no parameter server), for comparison.
require 'torch'
require 'nn'
torch.manualSeed(1234)
local model = nn.Sequential()
model:add(nn.Linear(2, 20)):add(nn.Tanh()):add(nn.Linear(20, 1))
local criterion = nn.MSECriterion()
local params, grads = model:getParameters()
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
    if step % 200 == 0 then print(string.format('step %d  loss %.4f', step, loss)) end
end
