"""TEST SCAFFOLDING — write the reference's generated model translation unit for a FlatModel.

The reference only knows how to generate `<model>_generated_model.cpp` from a `spatialpy.Model`
(spatialpy/solvers/solver.py:100-158).  The synthetic BASELINE configurations (spatialpy_b200/configs.py) are built
straight into arrays, so to run the UNMODIFIED reference engine on them — as parity oracle and as the CPU baseline —
this module fills the reference's own template (read from the checkout at build time, never copied into the repo)
with the same substitutions solver.py makes, field for field:

  __DEFINE_PARAMETERS__   solver.py:301-310     __INPUT_CONSTANTS__  solver.py:208-288
  __DEFINE_REACTIONS__    solver.py:344-373     __SYSTEM_CONFIG__    solver.py:375-419
  __DEFINE_CHEM_FUNS__    solver.py:160-190     __INIT_PARTICLES__   solver.py:312-331
  __DEFINE_GET_NEXT_OUTPUT__ solver.py:290-299  __BOUNDARY_CONDITIONS__ solver.py:131-133
"""

def _body(expr, restrict_to):
    if not restrict_to:
        return f"return {expr};"
    cond = "||".join(f"sd == {t}" for t in restrict_to)
    return f"if({cond}){{\nreturn {expr};\n}}else{{\n\treturn 0.0;}}"


def _arr(header, data, to_int=False):
    return f"{header}[{len(data)}] = {{" + ",".join(str(int(v)) if to_int else repr(float(v)) for v in data) + "};\n"


def emit(fm, path, template_path, debug_level=0):
    N, S, R = fm.num_particles, fm.num_species, fm.num_reactions
    Sc, Rc, Sd, Rd, ndf = fm.num_chem_species, fm.num_chem_rxns, fm.num_stoch_species, fm.num_stoch_rxns, fm.num_data_fn
    params = "".join(f"const double {k} = {float(v)!r};\n" for k, v in fm.parameters.items())
    params += "".join(f"const size_t {k} = {int(v)};\n" for k, v in fm.type_constants.items())
    funcs = funcinits = dets = detinits = ""
    for i, r in enumerate(fm.reactions):
        funcs += (f"double {r.name}(const int *x, double t, const double vol, const double *data_fn, int sd)\n{{\n"
                  f"{_body(r.propensity, r.restrict_to)}\n}}\n\n")
        funcinits += f"    ptr[{i}] = (PropensityFun) {r.name};\n"
        dets += (f"double det{r.name}(const double *x, double t, const double vol, const double *data_fn, int sd)\n{{\n"
                 f"{_body(r.ode_propensity, r.restrict_to)}\n}}\n\n")
        detinits += f"    ptr[{i}] = (ChemRxnFun) det{r.name};\n"
    # input constants
    const = f"unsigned int input_u0[{S * N}] = {{" + ",".join(str(int(v)) for v in fm.u0.reshape(-1)) + "};\n"
    if S > 0:
        if R > 0:
            const += _arr("static int input_N_dense", fm.N_dense.reshape(-1), True)
            const += _arr("static size_t input_irN", fm.irN, True)
            const += _arr("static size_t input_jcN", fm.jcN, True)
            const += _arr("static int input_prN", fm.prN, True)
            if ndf > 0:
                const += _arr("static double input_data_fn", fm.data_fn.reshape(-1))
        else:
            const += ("static int input_N_dense[0] = {};\nstatic size_t input_irN[0] = {};\n"
                      "static size_t input_jcN[0] = {};\nstatic int input_prN[0] = {};\n")
        const += _arr("static size_t input_irG", fm.irG, True)
        const += _arr("static size_t input_jcG", fm.jcG, True)
        const += "const char* const input_species_names[] = {" + ",".join(f'"{n}"' for n in fm.species_names) + ", 0};\n"
    const += f"const int input_num_subdomain = {fm.num_types};\n"
    const += _arr("const double input_subdomain_diffusion_matrix", fm.diffusion_matrix.reshape(-1))
    # system config
    cfg = f"debug_flag = {debug_level};\n"
    cfg += f"ParticleSystem *system = new ParticleSystem({fm.num_types},{Sc},{Rc},{Sd},{Rd},{ndf});\n"
    cfg += f"system->static_domain = {int(fm.static_domain)};\n"
    if S > 0:
        cfg += ("system->subdomain_diffusion_matrix = input_subdomain_diffusion_matrix;\n"
                "system->stoichiometric_matrix = input_N_dense;\n"
                "system->chem_rxn_rhs_functions = ALLOC_ChemRxnFun();\n"
                "system->stoch_rxn_propensity_functions = ALLOC_propensities();\n"
                "system->species_names = input_species_names;\n")
    cfg += f"system->dt = {fm.dt!r};\nsystem->nt = {int(fm.nt)};\nsystem->h = {fm.h!r};\n"
    cfg += f"system->rho0 = {fm.rho0!r};\nsystem->c0 = {fm.c0!r};\nsystem->P0 = {fm.P0!r};\n"
    cfg += f"system->xlo = {fm.xlim[0]!r};\nsystem->xhi = {fm.xlim[1]!r};\n"
    cfg += f"system->ylo = {fm.ylim[0]!r};\nsystem->yhi = {fm.ylim[1]!r};\n"
    cfg += f"system->zlo = {fm.zlim[0]!r};\nsystem->zhi = {fm.zlim[1]!r};\n"
    cfg += f"system->dimension = {int(fm.dimension)};\n"
    for i, g in enumerate(fm.gravity):
        cfg += f"system->gravity[{i}] = {float(g)!r};\n"
    inv_types = {}
    for k, v in fm.type_constants.items():
        inv_types.setdefault(int(v), k)
    parts = []
    X, NU, MA, CC, RH = fm.x.tolist(), fm.nu.tolist(), fm.mass.tolist(), fm.c.tolist(), fm.rho.tolist()   # python floats
    for i in range(N):
        t = int(fm.type[i])
        parts.append(f"init_create_particle(sys,id++,{X[i][0]!r},{X[i][1]!r},{X[i][2]!r},{inv_types.get(t, t)},"
                     f"{NU[i]!r},{MA[i]!r},{CC[i]!r},{RH[i]!r},{int(fm.solid[i])},{Sc});\n")
    nxt = ("unsigned int get_next_output(ParticleSystem* system)\n{\nstatic int index = 0;\n"
           "const std::vector<unsigned int> output_steps = {" + ", ".join(str(int(v)) for v in fm.output_steps) +
           "};\nunsigned int next_step = output_steps[index];\nindex++;\nreturn next_step;\n}\n")
    dfa = ""
    if S > 0:
        for k in range(ndf):
            dfa += f"this_particle->data_fn[{k}] = input_data_fn[{k}*{N}+id];"
    init_rdme = ("initialize_rdme(system, input_irN, input_jcN, input_prN, input_irG, input_jcG, input_u0);"
                 if (fm.enable_rdme and S > 0) else "")
    repl = {
        "__NUMBER_OF_REACTIONS__": str(R), "__NUMBER_OF_SPECIES__": str(S), "__NUMBER_OF_VOXELS__": str(N),
        "__DEFINE_PARAMETERS__": params, "__DEFINE_REACTIONS__": funcs, "__DEFINE_PROPFUNS__": funcinits,
        "__DEFINE_CHEM_FUNS__": dets, "__DEFINE_CHEM_FUN_INITS__": detinits,
        "__INIT_PARTICLES__": "".join(parts), "__DATA_FUNCTION_ASSIGN__": dfa, "__INPUT_CONSTANTS__": const,
        "__SYSTEM_CONFIG__": cfg, "__INIT_RDME__": init_rdme, "__BOUNDARY_CONDITIONS__": fm.bc_source or "",
        "__DEFINE_GET_NEXT_OUTPUT__": nxt,
    }
    with open(template_path, "r", encoding="utf-8") as f:
        text = f.read()
    for k, v in repl.items():     # same order as solver.py:135-153 (dict order) so nested tokens resolve identically
        text = text.replace(k, v)
    with open(path, "w", encoding="utf-8") as f:
        f.write(text)
    return path
