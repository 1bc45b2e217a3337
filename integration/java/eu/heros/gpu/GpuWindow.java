package eu.heros.gpu;

import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

import java.io.Serializable;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import nl.tudelft.simulation.medlabs.common.MedlabsRuntimeException;
import nl.tudelft.simulation.medlabs.model.MedlabsModelInterface;
import nl.tudelft.simulation.medlabs.simulation.TimeUnit;

/**
 * The per-window event of production mode: every {@code windowHours} of simulator time the stays that ended in the
 * window (collected by {@link GpuCovid19Transmission#onStayClosed}) and the stays still open (reported by the caller
 * through {@link OpenStays}) go to the engine in one heros_submit_visits call, heros_step carries the window out, and
 * the results come back: phase transitions are replayed on the Person objects
 * ({@link GpuCovid19Progression#applyTransitions}), infection events are handed to a listener (the place where the
 * reference's medlabs engine applies InfectionRecord.infectedPersons and feeds its monitors). Scheduled with the same
 * DSOL call the reference's Covid19Progression uses for its own events (Covid19Progression.java:192). Not compiled
 * here (see {@link HerosGpu}).
 */
public class GpuWindow implements Serializable
{
    /** */
    private static final long serialVersionUID = 1L;

    /** Walks the stays that are open at the end of a window: (person, location, sub-location, time of entry). */
    public interface OpenStays
    {
        /** @param sink StaySink; receives every person that is inside a location right now */
        void forEach(StaySink sink);
    }

    /** Receiver of {@link OpenStays#forEach}. */
    public interface StaySink
    {
        /** One open stay. */
        void accept(int personId, int locationId, short subIndex, double tIn);
    }

    /** Receiver of the infection events of a window, ordered by (time, person). */
    public interface InfectionListener
    {
        /** One person infected at a location at a time, with the probability of the evaluation that did it. */
        void infected(int personId, int locationId, short subIndex, double time, double pInfection);
    }

    private final MedlabsModelInterface model;

    private final HerosGpu gpu;

    private final GpuCovid19Transmission transmission;

    private final GpuCovid19Progression progression;

    private final OpenStays openStays;

    private final InfectionListener listener;

    private final double windowHours;

    /** Index of the next infection event of the engine's log that has not been delivered yet. */
    private long infectionCursor = 0;

    /**
     * @param model MedlabsModelInterface; the model whose simulator the window events are scheduled on
     * @param gpu HerosGpu; the engine of this replication
     * @param transmission GpuCovid19Transmission; switched to production mode here
     * @param progression GpuCovid19Progression; receives the transitions of every window
     * @param openStays OpenStays; the persons that are inside a location at the end of a window
     * @param listener InfectionListener; may be null
     * @param windowHours double; the window of the engine (HerosGpu constructor)
     */
    public GpuWindow(final MedlabsModelInterface model, final HerosGpu gpu, final GpuCovid19Transmission transmission,
            final GpuCovid19Progression progression, final OpenStays openStays, final InfectionListener listener,
            final double windowHours)
    {
        this.model = model;
        this.gpu = gpu;
        this.transmission = transmission;
        this.progression = progression;
        this.openStays = openStays;
        this.listener = listener;
        this.windowHours = windowHours;
        transmission.setWindowed(true);
        model.getSimulator().scheduleEventRel(windowHours, TimeUnit.HOUR, this, "endOfWindow", null);
    }

    /** The window that ends now. */
    protected void endOfWindow()
    {
        final double now = this.model.getSimulator().getSimulatorTime().doubleValue();
        final VisitBuffer visits = this.transmission.visits;
        // a stay still open goes with t_out = +inf and is submitted again, with the same t_in, once it has ended
        this.openStays.forEach((person, location, sub, tIn) -> visits.add(person, location, sub, tIn, Double.POSITIVE_INFINITY));
        try
        {
            this.gpu.check((int) HerosGpuNative.SUBMIT_VISITS.invokeExact(this.gpu.handle, visits.size(), visits.person(),
                    visits.location(), visits.sublocation(), visits.tIn(), visits.tOut()));
            visits.clear();
            // HEROS_ERR_INVALID "... submitted stays were rejected" = ids outside the tables or NaN times were sent; the
            // window has been carried out without them (the reference would have thrown on the unknown id)
            this.gpu.step(now);
            this.progression.applyTransitions();
            if (this.listener != null)
                deliverInfections();
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
        this.model.getSimulator().scheduleEventRel(this.windowHours, TimeUnit.HOUR, this, "endOfWindow", null);
    }

    private void deliverInfections() throws Throwable
    {
        final int chunk = 1 << 16;
        try (Arena a = Arena.ofConfined())
        {
            MemorySegment person = a.allocate(JAVA_INT, chunk), location = a.allocate(JAVA_INT, chunk);
            MemorySegment sub = a.allocate(JAVA_SHORT, chunk), time = a.allocate(JAVA_DOUBLE, chunk);
            MemorySegment p = a.allocate(JAVA_DOUBLE, chunk), nOut = a.allocate(JAVA_LONG);
            while (true)
            {
                int status = (int) HerosGpuNative.GET_INFECTIONS.invokeExact(this.gpu.handle, this.infectionCursor, (long) chunk,
                        person, location, sub, time, p, nOut);
                if (status != HerosGpu.ERR_CAPACITY)
                    this.gpu.check(status);
                long available = nOut.get(JAVA_LONG, 0);
                long n = Math.min(available, chunk);
                for (long i = 0; i < n; i++)
                    this.listener.infected(person.getAtIndex(JAVA_INT, i), location.getAtIndex(JAVA_INT, i),
                            sub.getAtIndex(JAVA_SHORT, i), time.getAtIndex(JAVA_DOUBLE, i), p.getAtIndex(JAVA_DOUBLE, i));
                this.infectionCursor += n;
                if (available <= chunk)
                    return;
            }
        }
    }
}
