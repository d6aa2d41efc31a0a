"""Optional chain figure of ``lightcurve_mcmc(show=..., save_plot_as=...)``.

Presentation only (SURVEY.md section 2 marks plotting out of scope): one column of walker traces per sampling phase,
one row per parameter.  Requires matplotlib, which is imported lazily so that the fitting path never depends on it.
"""


class ChainFigure:
    def __init__(self, labels):
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:                      # no silent skip: the caller asked for a figure
            raise ImportError('show / save_plot_as need matplotlib') from exc
        self.plt, self.labels = plt, list(labels)
        n = len(self.labels)
        self.fig, axes = plt.subplots(n, 2, figsize=(12., 2. * n), squeeze=False)
        self.columns = (axes[:, 0], axes[:, 1])

    def draw(self, column, chain, title):
        """``chain`` [nwalkers, nsteps, ndim] (``sampler.chain``): every walker's trace, one panel per parameter."""
        panels = self.columns[column]
        for i, (panel, label) in enumerate(zip(panels, self.labels)):
            panel.plot(chain[:, :, i].T, 'k', alpha=0.2)
            panel.set_ylabel(label)
            if column:
                panel.yaxis.set_label_position('right')
                panel.yaxis.tick_right()
        panels[0].set_title(title)
        panels[-1].set_xlabel('Step Number')

    def finish(self, save_as='', show=False):
        self.fig.tight_layout()
        if save_as:
            print('saving chain plot as ' + save_as)
            self.fig.savefig(save_as)
        if show:
            self.plt.show()
