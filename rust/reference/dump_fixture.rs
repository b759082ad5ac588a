// src/bin/dump_fixture.rs -- runs the reference's OWN solvers on the application's scene and writes fixtures the
// yasph2d_b200 repository's tests consume (tests/test_rust_fixtures.py).  NOT COMPILED where it was written.
//
//     cargo run --release --bin dump_fixture -- <dfsph|wcsph> <steps> <out_dir>
//
// Output (formats of yasph2d_b200/stateio.py):
//     rust_<solver>_scene.ysph          positions, velocities (zero), boundary as `reset_fluid` builds them (main.rs:177-196)
//     rust_<solver>_trajectory.jsonl    one JSON object per step: step, dt_ns (TimeManager::simulation_step after the step), kinetic_energy
//     rust_<solver>_step<k>.ysph        positions, velocities, densities, boundary after step k, k in {1, 10, 100, steps}
// Only public API of the crate is used (the solvers' iteration counts are private fields, dfsph.rs:24-32: they are pinned
// indirectly, through dt and the states).
use std::fs::File;
use std::io::{BufWriter, Write};
use std::time::Duration;

use ggez::graphics::Rect;
use yasph2d::sph::timemanager::{AdaptiveTimeStepTarget, SimulationStepConfig, TimeManager, TimerConfig};
use yasph2d::sph::{self, Solver};
use yasph2d::units::*;

// main.rs:177-196
fn reset_fluid(w: &mut sph::FluidParticleWorld) {
    w.remove_all_fluid_particles();
    w.remove_all_boundary_particles();
    w.add_fluid_rect(&Rect::new(0.1, 0.7, 0.5, 1.0), 0.05);
    w.add_boundary_thick_line(Point::new(0.0, 2.5), Point::new(2.0, 2.5), 4);
    w.add_boundary_thick_line(Point::new(0.0, 0.0), Point::new(2.0, 0.0), 4);
    w.add_boundary_thick_line(Point::new(0.0, 0.0), Point::new(0.0, 2.5), 4);
    w.add_boundary_thick_line(Point::new(2.0, 0.0), Point::new(2.0, 2.5), 4);
    w.add_boundary_thick_line(Point::new(0.0, 0.6), Point::new(1.75, 0.5), 2);
    w.add_boundary_thick_line(Point::new(0.0, 2.5), Point::new(2.0, 2.5), 2);
    w.add_boundary_thick_line(Point::new(-2.0, -0.5), Point::new(4.0, -0.5), 4);
}

fn flat2<T: Copy + Into<[f32; 2]>>(v: &[T]) -> Vec<f32> {
    v.iter().flat_map(|p| { let a: [f32; 2] = (*p).into(); a }).collect()
}

/// b"YSPH2D01" | u32 header bytes | u32 0 | JSON header padded to 8 | raw little-endian f32 arrays, each padded to 8 bytes
fn write_ysph(path: &str, params: &str, solver: &str, arrays: &[(&str, Vec<usize>, &[f32])]) -> std::io::Result<()> {
    let metas: Vec<String> = arrays
        .iter()
        .map(|(name, shape, _)| format!("{{\"name\": \"{}\", \"dtype\": \"f4\", \"shape\": [{}]}}", name, shape.iter().map(|s| s.to_string()).collect::<Vec<_>>().join(", ")))
        .collect();
    let mut header = format!("{{\"params\": {}, \"solver\": {}, \"arrays\": [{}]}}", params, solver, metas.join(", ")).into_bytes();
    while header.len() % 8 != 0 {
        header.push(b' ');
    }
    let mut f = BufWriter::new(File::create(path)?);
    f.write_all(b"YSPH2D01")?;
    f.write_all(&(header.len() as u32).to_le_bytes())?;
    f.write_all(&[0u8; 4])?;
    f.write_all(&header)?;
    for (_, _, data) in arrays {
        for x in data.iter() {
            f.write_all(&x.to_le_bytes())?;
        }
        if (data.len() * 4) % 8 != 0 {
            f.write_all(&[0u8; 4])?;
        }
    }
    Ok(())
}

fn dump_state(path: &str, w: &sph::FluidParticleWorld, params: &str, solver: &str) -> std::io::Result<()> {
    let p = &w.particles;
    let (pos, vel, bnd) = (flat2(&p.positions), flat2(&p.velocities), flat2(&p.boundary_particles));
    let n = p.positions.len();
    let mut arrays: Vec<(&str, Vec<usize>, &[f32])> = vec![("positions", vec![n, 2], &pos), ("velocities", vec![n, 2], &vel), ("boundary", vec![p.boundary_particles.len(), 2], &bnd)];
    if p.densities.len() == n {
        arrays.push(("densities", vec![n], &p.densities));
    }
    write_ysph(path, params, solver, &arrays)
}

fn main() -> std::io::Result<()> {
    let args: Vec<String> = std::env::args().collect();
    let (kind, steps, out) = (args[1].as_str(), args[2].parse::<usize>().unwrap(), args[3].as_str());
    let mut world = sph::FluidParticleWorld::new(2.0, 10000.0, 100.0); // main.rs:85-89
    reset_fluid(&mut world);
    world.particles.velocities.resize(world.particles.positions.len(), cgmath::Zero::zero());
    let xsph = sph::XSPHViscosityModel::new(world.properties.smoothing_length()); // main.rs:93
    let (mut solver, cfl): (Box<dyn Solver>, f32) = match kind {
        "wcsph" => (Box::new(sph::WCSPHSolver::new(xsph, &world.properties)), 0.2), // main.rs:99, 116
        _ => (Box::new(sph::DFSPHSolver::new(xsph, world.properties.smoothing_length())), 1.5), // main.rs:100, 117
    };
    let mut time = TimeManager::new(TimerConfig {
        step_config: SimulationStepConfig::AdaptiveTimeStep {
            timestep_max: Duration::from_secs_f32(1.0 / 120.0 / 3.0),  // main.rs:123
            timestep_min: Duration::from_secs_f32(1.0 / 60.0 / 400.0), // main.rs:124
            timestep_target_frame: AdaptiveTimeStepTarget::None,
            cfl_factor: cfl,
        },
        max_simulated_time_per_frame: Duration::from_secs_f64(1.0 / 30.0),
    });
    let params = format!("{{\"source\": \"rust reference\", \"solver\": \"{}\", \"cfl_factor\": {}, \"scene\": \"reset_fluid (main.rs:177-196)\"}}", kind, cfl);
    dump_state(&format!("{}/rust_{}_scene.ysph", out, kind), &world, &params, "{}")?;
    let mut traj = BufWriter::new(File::create(format!("{}/rust_{}_trajectory.jsonl", out, kind))?);
    writeln!(traj, "{{\"header\": {}, \"fields\": [\"step\", \"dt_ns\", \"kinetic_energy\"]}}", params)?;
    let mass = world.properties.particle_mass() as f64;
    for step in 1..=steps {
        solver.simulation_step(&mut world, &mut time); // main.rs:279
        let ekin: f64 = world.particles.velocities.iter().map(|v| 0.5 * mass * ((v.x as f64) * (v.x as f64) + (v.y as f64) * (v.y as f64))).sum();
        writeln!(traj, "{{\"step\": {}, \"dt_ns\": {}, \"kinetic_energy\": {:e}}}", step, time.simulation_step().as_nanos(), ekin)?;
        if step == 1 || step == 10 || step == 100 || step == steps {
            let s = format!("{{\"step\": {}, \"dt_ns\": {}}}", step, time.simulation_step().as_nanos());
            dump_state(&format!("{}/rust_{}_step{}.ysph", out, kind, step), &world, &params, &s)?;
        }
    }
    Ok(())
}
