"""Plain-text simulation log with the reference's line formats (simulation_utilities/sim_logger.py:23-93),
so that existing log parsers keep working.  On the device-resident path the per-step lines are
written after the fact from the statistics ring the step kernels fill (no per-step host sync)."""
from .Constants import Constants

__all__ = ['SimLogger']


def _wn(x):
    return Constants.convert(x, 'wavenumbers', to_AU=False)


class SimLogger:
    def __init__(self, fname, overwrite=False):
        self.fl = open(fname, 'w' if overwrite else 'a')

    # ---- run level
    def write_beginning(self, attribs):
        w = self.fl.write
        w(f"Simulation {attribs['sim_name']} starting at step {attribs['cur_timestep']}\n")
        w("Potential attributes: \n")
        for key, value in attribs['potential_info'].items():
            if not key.startswith("_"):
                w(f"\t{key}: {value}\n")
        if attribs['impsamp_manager'] is not None:
            w("Imp Samp attributes: \n")
            for key, value in attribs['imp_info'].items():
                if not key.startswith("_"):
                    w(f"\t{key}: {value}\n")
        w(f"Num Walkers: {attribs['num_walkers']}\n")
        w(f"Num Time Steps: {attribs['num_timesteps']}\n")
        w(f"Weighting Type: {attribs['weighting']}\n")
        w(f"Branch every {attribs['branch_every']} Time Step(s)\n")
        w(f"Delta Tau: {attribs['delta_t']} a.u.\n")
        w(f"Start Structure Array Shape: {attribs['start_structures'].shape}\n")
        w(f"Masses of Each Atom: {attribs['masses']}\n")
        w(f"Equilibration Steps Before Collecting Wave Functions: {attribs['equil_steps']}\n")
        w(f"Checkpoint Every {attribs['chkpt_every']} time steps\n")
        w(f"Collect Wave Functions Every {attribs['wfn_every']} Time Steps After Equilibration\n")
        w("\n")

    def final_chkpt(self):
        self.fl.write("\nFinal checkpoint is written to chkpts folder.\n\n")

    def finish_sim(self, final_time):
        self.fl.write("Simulation has finished.\n\n")
        self.fl.write(f"Simulation took {final_time} seconds.\n\n")
        self.fl.close()

    # ---- step level
    def write_ts(self, cur_time_step):
        self.fl.write(f"Time step {cur_time_step}\n")

    def write_chkpt(self, cur_time_step):
        self.fl.write(f"Checkpointing, time step {cur_time_step}\n")

    def write_wfn_save(self, cur_time_step):
        self.fl.write(f"Starting descendant weighting, time step {cur_time_step}\n")
        self.fl.write(f"Will save wave function from time step {cur_time_step}\n")

    def write_desc_wt(self, cur_time_step):
        self.fl.write(f"Finished descendant weighting at end of time step {cur_time_step}\n")
        self.fl.write("Saving wave function with descendant weights\n")

    def write_pot_time(self, cur_time_step, pot_time, maxpot, minpot, avgpot):
        w = self.fl.write
        w(f"Potential call time at time step {cur_time_step}:\n")
        w(f"\t{pot_time} seconds\n")
        w(f"\tAverage energy of ensemble: {_wn(avgpot)} cm-1 (without vref correction)\n")
        w(f"\tHighest energy walker: {_wn(maxpot)} cm-1\n")
        w(f"\tLowest energy walker: {_wn(minpot)} cm-1\n")

    def write_local(self, local_energy):
        self.fl.write(f"\tAverage local energy of ensemble: {_wn(local_energy)} cm-1 \n")

    def write_branching(self, cur_time_step, weighting, birthdeath_branch):
        w = self.fl.write
        if weighting == 'discrete':
            w(f"Birth/Death at time step {cur_time_step}:\n")
            w(f"\tWalker Births: {birthdeath_branch[0]}\n")
            w(f"\tWalker Deaths: {birthdeath_branch[1]}\n")
            w(f"\tNumber of Walkers: {birthdeath_branch[2]}\n")
        elif weighting == 'continuous':
            w(f"Branching at time step {cur_time_step}:\n")
            w(f"\tWalkers Branched: {birthdeath_branch[0]}:\n")
            w(f"\tMax Wt before Branched: {birthdeath_branch[1]}:\n")
            w(f"\tMin Wt before Branched: {birthdeath_branch[2]}:\n")

    def write_rejections(self, rejected, total):
        self.fl.write(f"Metropolis rejected {rejected} of {total} walkers ({(rejected / total) * 100:0.2f} %)\n")

    def write_imp_disp_time(self, time):
        self.fl.write(f"Imp samp displacement took {time} seconds.\n")
