#!/bin/sh
# Build the managed assembly and pack it with the native library (counterpart of the
# reference's GenerateNugetPackage.ps1). Needs the .NET SDK and nuget on PATH.
set -e
cd "$(dirname "$0")/.."
dotnet build -c Release MultiversoCLR.csproj
nuget pack NuGet/MultiversoCLR.nuspec -OutputDirectory NuGet
