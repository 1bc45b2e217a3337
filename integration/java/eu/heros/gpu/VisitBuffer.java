package eu.heros.gpu;

import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

/**
 * The stays of one window as five off-heap columns -- the argument layout of heros_submit_visits(person, location,
 * sublocation, t_in, t_out). Grows by doubling; cleared (not freed) after every window. Not compiled here (see
 * {@link HerosGpu}).
 */
final class VisitBuffer
{
    private Arena arena = Arena.ofConfined();

    private MemorySegment person, location, sublocation, tIn, tOut;

    private long capacity, size;

    VisitBuffer(final long initialCapacity)
    {
        allocate(initialCapacity);
    }

    private void allocate(final long n)
    {
        this.person = this.arena.allocate(JAVA_INT, n);
        this.location = this.arena.allocate(JAVA_INT, n);
        this.sublocation = this.arena.allocate(JAVA_SHORT, n);
        this.tIn = this.arena.allocate(JAVA_DOUBLE, n);
        this.tOut = this.arena.allocate(JAVA_DOUBLE, n);
        this.capacity = n;
    }

    private void grow()
    {
        Arena old = this.arena;
        MemorySegment p = this.person, l = this.location, s = this.sublocation, i = this.tIn, o = this.tOut;
        this.arena = Arena.ofConfined();
        allocate(2 * this.capacity);
        MemorySegment.copy(p, 0, this.person, 0, 4 * this.size);
        MemorySegment.copy(l, 0, this.location, 0, 4 * this.size);
        MemorySegment.copy(s, 0, this.sublocation, 0, 2 * this.size);
        MemorySegment.copy(i, 0, this.tIn, 0, 8 * this.size);
        MemorySegment.copy(o, 0, this.tOut, 0, 8 * this.size);
        old.close();
    }

    /** One stay; tOut = +infinity while the person is still there at the end of the window. */
    void add(final int personId, final int locationId, final short subIndex, final double in, final double out)
    {
        if (this.size == this.capacity)
            grow();
        this.person.setAtIndex(JAVA_INT, this.size, personId);
        this.location.setAtIndex(JAVA_INT, this.size, locationId);
        this.sublocation.setAtIndex(JAVA_SHORT, this.size, subIndex);
        this.tIn.setAtIndex(JAVA_DOUBLE, this.size, in);
        this.tOut.setAtIndex(JAVA_DOUBLE, this.size, out);
        this.size++;
    }

    void clear()
    {
        this.size = 0;
    }

    long size()
    {
        return this.size;
    }

    MemorySegment person()
    {
        return this.person;
    }

    MemorySegment location()
    {
        return this.location;
    }

    MemorySegment sublocation()
    {
        return this.sublocation;
    }

    MemorySegment tIn()
    {
        return this.tIn;
    }

    MemorySegment tOut()
    {
        return this.tOut;
    }
}
