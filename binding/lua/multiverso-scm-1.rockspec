package = "multiverso"
version = "scm-1"
source = { url = "git://github.com/multiverso-b200/multiverso-b200" }
description = {
   summary = "Lua/Torch binding of multiverso-b200",
   detailed = "LuaJIT FFI binding over the C API of libmultiverso.so (API compatible with Microsoft/multiverso's torch binding).",
   license = "MIT"
}
dependencies = { "lua >= 5.1", "torch >= 7.0" }
build = {
   type = "builtin",
   modules = {
      ["multiverso.init"] = "init.lua",
      ["multiverso.util"] = "util.lua",
      ["multiverso.ArrayTableHandler"] = "ArrayTableHandler.lua",
      ["multiverso.MatrixTableHandler"] = "MatrixTableHandler.lua",
   }
}
