// lj_deck.cpp -- the reference's LJ regression deck (contribs/microStamp/samples/benchmark_lj_snap/input_lj_Ni.msp) composed
// from the operator mirror exactly as the YAML batches of data/config/*.msp compose it:
//   input_data -> nbh_dist -> first_iteration -> compute_loop(numerical_scheme) -> simulation_epilog(check_values)
// usage: lj_deck <check_values file> [grid cells per axis = 4] [steps = 100] [check | dump <prefix> | restart <dump file>]
//   dump: after the loop write <prefix>.dump (write_dump) and <prefix>.xyz (write_xyz), then check_values
//   restart: input_data = read_dump instead of lattice + noise, then the loop, then check_values
#include "xnb_operators.hpp"
#include <cstdio>
#include <cstdlib>
#include <string>

using namespace xnb::host;

static const char* dir_name(SlotDirection d) { return d == INPUT ? "INPUT" : d == OUTPUT ? "OUTPUT" : d == INPUT_OUTPUT ? "INPUT_OUTPUT" : "PRIVATE"; }

int main(int argc, char** argv)
{
  if (argc > 1 && std::string(argv[1]) == "--list")
  {
    // operator names and slots, one per line: "operator slot DIRECTION [REQUIRED]"
    register_hot_path_operators();
    for (const std::string& n : OperatorNodeFactory::instance()->available_operators())
    {
      auto op = OperatorNodeFactory::instance()->make_operator(n);
      std::printf("%s\n", n.c_str());
      for (SlotBase* s : op->slots) std::printf("%s %s %s%s\n", n.c_str(), s->name.c_str(), dir_name(s->dir), s->required ? " REQUIRED" : "");
    }
    return 0;
  }
  const std::string golden = argc > 1 ? argv[1] : "check_values_lj_Ni.dat";
  const int cells = argc > 2 ? std::atoi(argv[2]) : 4;
  const int end_iteration = argc > 3 ? std::atoi(argv[3]) : 100;
  const std::string mode = argc > 4 ? argv[4] : "";
  const std::string io_arg = argc > 5 ? argv[5] : "";
  char dom[512];
  std::snprintf(dom, sizeof dom, "{ cell_size: 13.92 ang , grid_dims: [ %d , %d , %d ] , bounds: [ [ 0.0 um , 0.0 um , 0.0 um ] , [ %.2f ang , %.2f ang , %.2f ang ] ] , "
                "periodic: [ true , true , true ] , expandable: false }", cells, cells, cells, 13.92 * cells, 13.92 * cells, 13.92 * cells);

  Batch sim;
  // global: (input_lj_Ni.msp:36-43)
  *sim.value<double>("dt") = convert_quantity("2e-3 ps");
  *sim.value<double>("rcut_inc") = convert_quantity("2.0 ang");
  *sim.value<double>("rcut_max") = convert_quantity("4.1 ang");
  *sim.value<bool>("deterministic_noise") = true;

  // input_data: (input_lj_Ni.msp:57-76)
  auto input_data = sim.sub_batch();
  input_data->add("particle_type_add_properties", "{ Ni: { mass: 58.693 Da , z: 28 } }");
  if (mode == "restart")
  {
    input_data->add("read_dump", "{ filename: " + io_arg + " }");      // domain and particles come from the checkpoint
    input_data->add("domain");
    input_data->add("init_rcb_grid");
  }
  else
  {
    input_data->add("domain", dom);
    input_data->add("init_rcb_grid");
    input_data->add("lattice", "{ structure: FCC , types: [ Ni , Ni , Ni , Ni ] , size: [ 3.48 ang , 3.48 ang , 3.48 ang ] }");
    input_data->add("gaussian_noise_r", "{ sigma: 0.05 ang }");
  }

  // compute_all_forces_energy: (input_lj_Ni.msp:88-92); lennard_jones_force: (:83-85)
  auto forces = sim.sub_batch();
  forces->add("zero_particle_force", "{ ghost: true }");
  forces->add("lennard_jones_force", "{ config: { epsilon: 0.3729 eV , sigma: 2.2808 ang } , rcut: 4.1 ang }");
  forces->add("update_force_from_ghost");
  forces->add("divide_force_by_type_scalar");

  auto nbh_dist = sim.sub_batch();
  nbh_dist->add("nbh_dist");

  // parallel_update_particles + update_particle_neighbors (update-particles.msp:42-53)
  auto add_parallel_update = [](Batch& b)
  {
    b.add("migrate_cell_particles"); b.add("rebuild_amr", "{ sub_grid_density: 6.5 }"); b.add("backup_r"); b.add("ghost_comm_scheme"); b.add("ghost_update_all");
    b.add("amr_grid_pairs"); b.add("chunk_neighbors", "{ config: { build_particle_offset: true , chunk_size: 1 } }"); b.add("resize_particle_locks");
  };
  auto init_particles = sim.sub_batch();            // update-particles.msp:55-60 (static blocks: no load balance)
  init_particles->add("move_particles"); add_parallel_update(*init_particles);
  auto update_particles_full = sim.sub_batch();     // :62-68
  update_particles_full->add("move_particles"); add_parallel_update(*update_particles_full);
  auto update_particles_fast = sim.sub_batch();     // :70-73
  update_particles_fast->add("ghost_update_r");
  auto trigger_move_particles = sim.sub_batch();    // compute-loop.msp: rebind { threshold: max_displ , result: trigger_move_particles }
  trigger_move_particles->add("particle_displ_over", "", {{"threshold", "max_displ"}, {"result", "trigger_move_particles"}});
  auto verlet_first_half = sim.sub_batch();         // numerical-scheme.msp:13-15
  verlet_first_half->add("push_f_v_r", "{ dt_scale: 1.0 }"); verlet_first_half->add("push_f_v", "{ dt_scale: 0.5 }");
  auto verlet_second_half = sim.sub_batch();        // :17-18
  verlet_second_half->add("push_f_v", "{ dt_scale: 0.5 }");
  auto epilog = sim.sub_batch();                    // input_lj_Ni.msp:96-102
  epilog->add("check_values", "{ file: " + golden + " , pos_threshold: 1.e-5 , vel_threshold: 1.e-5 , acc_threshold: 1.e-5 }");

  // default_simulation (main-config.msp:33-49)
  input_data->execute();
  nbh_dist->execute();
  init_particles->execute(); forces->execute();     // first_iteration (compute-loop.msp:1-7)
  int rebuilds = 0;
  for (int it = 0; it < end_iteration; it++)        // compute_loop / numerical_scheme (numerical-scheme.msp:21-25)
  {
    verlet_first_half->execute();
    trigger_move_particles->execute();              // check_and_update_particles (update-particles.msp:75-78)
    if (*sim.value<bool>("trigger_move_particles")) { update_particles_full->execute(); rebuilds++; } else update_particles_fast->execute();
    forces->execute();
    verlet_second_half->execute();
  }
  if (mode == "dump")
  {
    auto io = sim.sub_batch();
    io->add("write_dump", "{ filename: " + io_arg + ".dump }");
    io->add("write_xyz", "{ filename: " + io_arg + ".xyz , fields: [ velocity , id , type ] }");
    io->execute();
  }
  // the reference's golden file belongs to the verbatim deck; any other size runs check_values only on request (it then writes or reads its own file)
  if ((cells == 4 && end_iteration == 100) || !mode.empty()) epilog->execute();
  std::printf("lj_deck: %lld particles, %d iterations, %d neighbour rebuilds\n", (long long)sim.value<Grid>("grid")->number_of_particles(), end_iteration, rebuilds);
  return 0;
}
