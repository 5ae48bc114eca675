# parity_dump.jl -- for whoever has Julia + LinearMPC.jl: dump the REAL reference's answers on the
# seeded theta set so the oracle's unpinned quantities (final active set, exit flag, iterations vs
# libdaqp) can be checked.  Not runnable in the build container (no Julia).
#   julia parity_dump.jl invpend 10000 out.csv
using LinearMPC, Random, DelimitedFiles
name, N, out = ARGS[1], parse(Int, ARGS[2]), ARGS[3]
mpc, range = LinearMPC.mpc_examples(name)
LinearMPC.setup!(mpc)
lo = [range.xmin; range.rmin; range.dmin[1:mpc.nd]; range.umin[1:mpc.nuprev]]
hi = [range.xmax; range.rmax; range.dmax[1:mpc.nd]; range.umax[1:mpc.nuprev]]
Random.seed!(1234)
rows = Matrix{Float64}(undef, N, length(lo) + mpc.model.nu + 2 + length(mpc.mpQP.bu))
for q in 1:N
    th = lo .+ (hi .- lo) .* rand(length(lo))
    z, fval, flag, info = LinearMPC.solve(mpc, th)          # cold/warm as the Julia path does
    u = z[1:mpc.model.nu] - mpc.K * th[1:mpc.model.nx]
    rows[q, :] = [th; u; flag; info.iterations; sign.(info.λ)]
end
writedlm(out, rows, ',')
