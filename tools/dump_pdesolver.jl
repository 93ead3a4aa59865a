# Dump of one evalResidual call of PDESolver.jl (Julia 0.6 syntax, as the reference) in the flat PDSDUMP1 format that
# pdesolver.jl_b200/dump.py reads (SURVEY.md §8(c) "Consequence"): operator, mesh arrays, options, q and the residual Julia
# computed.  Run wherever the reference runs:
#
#     julia tools/dump_pdesolver.jl input_vals_3d_rk4.jl dump_3d.pds
#
# then, on the B200 side:   python -m pytest tests/test_dump.py --dump dump_3d.pds     (oracle AND CUDA path vs Julia's res)
using PDESolver, EulerEquationMod, ODLCommonTools

function wrec(io, name::String, a::Array{Float64})
  write(io, Int32(length(name))); write(io, name)
  write(io, Int32(1)); write(io, Int32(ndims(a)))
  for d in size(a); write(io, Int64(d)); end
  write(io, a)
end
function wrec(io, name::String, a::Array{Int64})
  write(io, Int32(length(name))); write(io, name)
  write(io, Int32(2)); write(io, Int32(ndims(a)))
  for d in size(a); write(io, Int64(d)); end
  write(io, a)
end
function wrec(io, name::String, s::String)
  b = Vector{UInt8}(s)
  write(io, Int32(length(name))); write(io, name)
  write(io, Int32(3)); write(io, Int32(1)); write(io, Int64(length(b))); write(io, b)
end

function dump_case(input_file::String, out::String)
  mesh, sbp, eqn, opts = createObjects(input_file)      # src/startup_func.jl; applies the IC named in the input file
  t = 0.0
  evalResidual(mesh, sbp, eqn, opts, t)                 # src/solver/euler/euler.jl:111-175
  f = mesh.sbpface
  sparse = isa(f, SummationByParts.SparseFace)
  itf = zeros(Int64, 5, mesh.numInterfaces)
  for (i, I) in enumerate(mesh.interfaces)
    itf[:, i] = [I.elementL, I.elementR, I.faceL, I.faceR, I.orient]
  end
  bf = zeros(Int64, 2, mesh.numBoundaryFaces)
  for (i, B) in enumerate(mesh.bndryfaces)
    bf[:, i] = [B.element, B.face]
  end
  keys_out = ["Flux_name", "Volume_flux_name", "SRCname", "volume_integral_type", "face_integral_type",
              "FaceElementIntegral_name", "gamma", "R", "Ma", "aoa", "p_free", "T_free", "operator_type", "order"]
  for i = 1:opts["numBC"]; push!(keys_out, "BC$(i)_name"); end
  lines = String[]
  for k in keys_out
    haskey(opts, k) && push!(lines, "$k=$(opts[k])")
  end
  open(out, "w") do io
    write(io, "PDSDUMP1"); write(io, Int32(23))
    wrec(io, "dim", Int64[mesh.dim]); wrec(io, "numDofPerNode", Int64[mesh.numDofPerNode])
    wrec(io, "degree", Int64[sbp.degree])
    wrec(io, "Q", Array{Float64}(sbp.Q)); wrec(io, "w", Array{Float64}(sbp.w))
    wrec(io, "interp", sparse ? ones(Float64, 1, f.numnodes) : Array{Float64}(f.interp))
    wrec(io, "perm", Array{Int64}(f.perm)); wrec(io, "nbrperm", Array{Int64}(f.nbrperm))
    wrec(io, "wface", Array{Float64}(f.wface)); wrec(io, "sparse_face", Int64[sparse ? 1 : 0])
    wrec(io, "coords", Array{Float64}(mesh.coords)); wrec(io, "dxidx", Array{Float64}(mesh.dxidx))
    wrec(io, "jac", Array{Float64}(mesh.jac)); wrec(io, "nrm_face", Array{Float64}(mesh.nrm_face))
    wrec(io, "nrm_bndry", Array{Float64}(mesh.nrm_bndry)); wrec(io, "coords_bndry", Array{Float64}(mesh.coords_bndry))
    wrec(io, "interfaces", itf); wrec(io, "bndryfaces", bf)
    wrec(io, "bndry_offsets", Array{Int64}(mesh.bndry_offsets))
    wrec(io, "opts", join(lines, "\n"))
    wrec(io, "q", Array{Float64}(eqn.q)); wrec(io, "res", Array{Float64}(eqn.res))
    wrec(io, "t", Float64[t])
  end
end

dump_case(ARGS[1], ARGS[2])
